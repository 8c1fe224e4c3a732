# Round-2 GPU call 1:  gpurun --timeout 2700 -- 'bash tools/r02_gpu1.sh'
# (a) HEAD (keyed seed stage) vs the library built before the single-writer change (old scalar seed kernel), seed-warp kernel, sort / occupancy / split sweeps (chr21-sized genome, 1 M pairs)
# (b) launch list + ncu --set full of the hot kernels at HEAD
# (c) the new config-size gate tests
# (d) first GRCh38-sized index build + bench
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/g1_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g1_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/seed \1 stitch \2 total \3 ms_heavy \4/')"; }
run base A=1
el base done
run pre_single_writer STAR_B200_LIB=$PWD/star_b200/lib_ab/libstar_b200_pre_single_writer.so
for c in 8; do run seedwarp$c STAR_B200_SEED_WARP=$c; done
run sort0 STAR_B200_SEED_SORT_BITS=0
run sort28 STAR_B200_SEED_SORT_BITS=28
run keyed4 STAR_B200_SEED_KEYED_CTAS_PER_SM=4
run keyed12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
for s in 20 40; do run split$s STAR_B200_HEAVY_SPLIT=$s; done
run dfs6 STAR_B200_FLAT_DFS_CTAS_PER_SM=6
run setup4 STAR_B200_FLAT_SETUP_CTAS_PER_SM=4
el sweeps done
# seed-warp parity on the chr21 workload (20 k pairs)
python - <<'PY' > gpurun_out/g1_seedwarp_parity.log 2>&1
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench, synth, oracle_capi as oc, star_b200 as sb
wd = "/tmp/star_b200_bench/chr21"
chrs, trs, idx, _ = bench.prepare_genome(wd, "chr21")
lib = sb.load_library(); index = sb.Index(lib, idx); ol = oc.load_oracle()
for rl, mm, n in ((100, 0.005, 20000), (150, 0.05, 5000)):
    m1, m2 = synth.make_reads(chrs, trs, n, read_len=rl, mm=mm, seed=5)
    seq, off, n_, nm = sb.pack_reads([m1, m2])
    oe = oc.OracleEngine(ol, index); res_o, al_o, st_o = oe.map_chunk(seq, off, n_, nm); oe.close()
    for ctas in ("0", "8"):
        os.environ["STAR_B200_SEED_WARP"] = ctas
        eng = sb.Engine(lib, index, max_reads=n_); res_g, al_g, st_g = eng.map_chunk(seq, off, n_, nm); eng.close()
        d = oc.compare_outputs(res_o, al_o, res_g, al_g)
        print("SEED_WARP", ctas, rl, mm, "diffs", len(d), "searches equal", st_g.mmp_searches == st_o.mmp_searches, d[:2], flush=True)
PY
tail -4 gpurun_out/g1_seedwarp_parity.log
el parity done
# launch list + full captures of the hot kernels (262144 pairs keep the replays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g1_launches.csv python bench.py --preset chr21 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g1_launch_bench.log 2>&1
for k in flat_dfs_warp_kernel flat_setup_kernel flat_record_warp_kernel seed_keyed_search_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/g1_$k python bench.py --preset chr21 --pairs 262144 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g1_ncu_$k.log 2>&1
  ncu -i gpurun_out/g1_$k.ncu-rep --page raw --csv > gpurun_out/g1_${k}_raw.csv 2>/dev/null
done
el ncu done
timeout 1500 python -m pytest tests/test_gpu_config_gate.py -m gpu -x -q > gpurun_out/g1_gate_tests.log 2>&1; tail -5 gpurun_out/g1_gate_tests.log
el gate tests done
# chr21 bench line at HEAD (with the command-line leg and the reference)
timeout 900 python bench.py --preset chr21 --steps 5 --warmup 3 > gpurun_out/g1_bench_chr21.json 2> gpurun_out/g1_bench_chr21.log; tail -c 1500 gpurun_out/g1_bench_chr21.json
el chr21 bench done
# GRCh38-sized: index build + bench
free -g | head -2; nproc; df -h /tmp | tail -1
STAR_B200_SA_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 3 --warmup 2 > gpurun_out/g1_bench_grch38.json 2> gpurun_out/g1_bench_grch38.log
echo "grch38 rc=$?"; tail -5 gpurun_out/g1_bench_grch38.log; tail -c 2500 gpurun_out/g1_bench_grch38.json
cp /tmp/star_b200_bench/grch38/gen_Log.out gpurun_out/g1_grch38_gen_Log.out 2>/dev/null
ls -la /tmp/star_b200_bench/grch38/idx 2>/dev/null | head -20
el all done
