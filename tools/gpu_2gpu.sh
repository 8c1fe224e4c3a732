mkdir -p gpurun_out /tmp/g2
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 600 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
cd /tmp/g2 && tar xzf $GRAFT_REPO_ROOT/tests/golden/tiny.tar.gz && ls | head -3; G=$(dirname $(find /tmp/g2 -name std_1.fq | head -1)); cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 -m star_b200.dist -- --genomeDir $G/idx --readFilesIn $G/std_1.fq $G/std_2.fq --outFileNamePrefix /tmp/g2/out/ --runThreadN 4 > gpurun_out/dist2.log 2>&1; echo rc=$?
python - <<PY
import sys,os
sys.path.insert(0,'tests'); import conftest as cf
G="$G"
a=cf.sam_body('/tmp/g2/out/Aligned.out.sam'); b=cf.sam_body(os.path.join(G,'ref_std','Aligned.out.sam'))
print('SAM identical', a==b, len(a)); print('SJ identical', open('/tmp/g2/out/SJ.out.tab','rb').read()==open(os.path.join(G,'ref_std','SJ.out.tab'),'rb').read())
print('Log counters identical', cf.log_counters('/tmp/g2/out/Log.final.out')==cf.log_counters(os.path.join(G,'ref_std','Log.final.out')))
PY
