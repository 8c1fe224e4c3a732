# Round-2 first GPU call (prepared at the end of round 1, when the GPU budget was spent).  Run with:
#   gpurun --timeout 1500 -- 'bash tools/r02_experiments.sh'
# 1. warp-uniform seed kernel (seed_warp.cuh, validated on the CPU emulation only so far): outputs vs oracle, then timing
# 2. prefix-split threshold of the task kernel: 82 % of the recursion nodes sit in windows with >= 15 seeds
# 3. launch list of the default configuration
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r02_seedwarp_parity.log 2>&1
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import tarfile, tempfile
import conftest as cf, oracle_capi as oc, star_b200 as sb
d = tempfile.mkdtemp(); tarfile.open("tests/golden/tiny.tar.gz").extractall(d); g = os.path.join(d, "tiny")
lib = sb.load_library(); idx = sb.Index(lib, os.path.join(g, "idx")); ol = oc.load_oracle()
for ctas in ("6", "8", "12"):
    os.environ["STAR_B200_SEED_WARP"] = ctas
    for name, files in (("std", ["std_1.fq", "std_2.fq"]), ("hard", ["hard_1.fq", "hard_2.fq"]), ("se", ["se_1.fq"])):
        mates = [cf.read_fastq_seqs(os.path.join(g, f)) for f in files]
        seq, off, n, nm = sb.pack_reads(mates)
        oe = oc.OracleEngine(ol, idx); res_o, al_o, st_o = oe.map_chunk(seq, off, n, nm); oe.close()
        eng = sb.Engine(lib, idx, max_reads=n); res_g, al_g, st_g = eng.map_chunk(seq, off, n, nm); eng.close()
        diffs = oc.compare_outputs(res_o, al_o, res_g, al_g)
        print("SEED_WARP", ctas, name, "diffs", len(diffs), "searches equal", st_g.mmp_searches == st_o.mmp_searches, "sai equal", st_g.mmp_sai_words == st_o.mmp_sai_words, diffs[:3], flush=True)
PY
cat gpurun_out/r02_seedwarp_parity.log | tail -12
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/r02_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/r02_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/seed \1 stitch \2 total \3 ms_heavy \4/') | $(grep 'heavy kernel warp' gpurun_out/r02_$tag.log | tail -1 | cut -c1-110)"; }
run base A=1
# A/B of the last kernel change of round 1 (single-writer discipline of the flat kernels, never timed): star_b200/lib_ab/ holds the library
# built from the commit before it (same ABI for this tool); built by `git worktree add build/wt 83ba07a^ && make` in the build container
if [ -e star_b200/lib_ab/libstar_b200_pre_single_writer.so ]; then run pre_single_writer STAR_B200_LIB=$PWD/star_b200/lib_ab/libstar_b200_pre_single_writer.so; fi
for c in 6 8 12; do run seedwarp$c STAR_B200_SEED_WARP=$c; done
for s in 16 20 30 52; do run split$s STAR_B200_HEAVY_SPLIT=$s; done
# 4. the round-1 kernels that were written after the GPU budget was spent and have only run under the host emulation:
#    GPU test-suite of junction insertion / 2-pass / genomeGenerate, then a timing of the suffix-array build and of an insertion on a
#    48 Mb random genome (chr21 size), with the launch list of the genomeGenerate run
timeout 900 python -m pytest tests/test_zz_twopass.py tests/test_zz_genome_generate.py -m gpu -q > gpurun_out/r02_new_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_new_gpu_tests.log
python - <<'PY'
import numpy as np
rng = np.random.default_rng(5)
with open("/tmp/g48.fa", "w") as f:
    for c in range(3):
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 16_000_000)].tobytes().decode()
        f.write(">chr%d\n" % (c + 1))
        for i in range(0, len(s), 80):
            f.write(s[i:i + 80] + "\n")
with open("/tmp/sj48.tab", "w") as f:   # 20 k random junctions
    for i in range(20000):
        a = int(rng.integers(1000, 15_900_000)); L = int(rng.integers(50, 20000))
        f.write("chr%d\t%d\t%d\t%s\n" % (1 + i % 3, a, a + L, "+-"[i % 2]))
PY
( time star_b200/bin/STAR --runMode genomeGenerate --genomeDir /tmp/idx48 --genomeFastaFiles /tmp/g48.fa --genomeSAindexNbases 12 --outFileNamePrefix /tmp/gen48_ ) > gpurun_out/r02_generate48.log 2>&1; tail -8 gpurun_out/r02_generate48.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_generate48_launches.csv star_b200/bin/STAR --runMode genomeGenerate --genomeDir /tmp/idx48b --genomeFastaFiles /tmp/g48.fa --genomeSAindexNbases 12 --outFileNamePrefix /tmp/gen48b_ > /dev/null 2>&1
head -c 400 tests/golden/tiny.tar.gz > /dev/null
python - <<'PY'
# a 1-read FASTQ is enough to time the insertion itself
open("/tmp/one.fq", "w").write("@r1\n" + "ACGT" * 25 + "\n+\n" + "I" * 100 + "\n")
PY
( time star_b200/bin/STAR --genomeDir /tmp/idx48 --readFilesIn /tmp/one.fq --sjdbFileChrStartEnd /tmp/sj48.tab --sjdbOverhang 99 --outFileNamePrefix /tmp/ins48_ ) > gpurun_out/r02_insert48.log 2>&1; tail -6 gpurun_out/r02_insert48.log; grep -i "SA search\|inserting\|Finished" /tmp/ins48_Log.out | tail -8 >> gpurun_out/r02_insert48.log
