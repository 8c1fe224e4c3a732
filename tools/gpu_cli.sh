mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cli" 2>&1 | tail -2
python - <<'PY'
import os,sys,time,subprocess
sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import bench, synth
wd='/tmp/star_b200_bench/chr21'; os.makedirs(wd,exist_ok=True)
chrs,trs,idx=bench.prepare_genome(wd,'chr21')
n=1000000
m1,m2=synth.make_reads(chrs,trs,n,read_len=100,mm=0.005,seed=77)
synth.write_fastq(m1,wd+'/cli_1.fq'); synth.write_fastq(m2,wd+'/cli_2.fq')
for thr in (32,):
    out=wd+'/cli_out_%d/'%thr
    t0=time.time()
    r=subprocess.run(['star_b200/bin/STAR','--runMode','alignReads','--genomeDir',idx,'--readFilesIn',wd+'/cli_1.fq',wd+'/cli_2.fq','--outFileNamePrefix',out,'--runThreadN',str(thr)],capture_output=True,text=True)
    dt=time.time()-t0
    log=open(out+'Log.out').read()
    eng=[l for l in log.split('\n') if 'star-b200' in l]
    print('threads',thr,'rc',r.returncode,'wall %.2f s'%dt,'pairs/s %.0f'%(n/dt),eng, flush=True)
    print(r.stdout[-400:])
PY
