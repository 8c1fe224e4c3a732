#!/usr/bin/env python3
"""Product-path multi-GPU measurement (BASELINE.json configs[3] / configs[4] shape): `python -m star_b200.dist` over ONE shared pair of
FASTQ files (strong scaling: the same input at every N), plain mapping and --twopassMode Basic.

  python tools/product_scale.py prepare [--preset chr21|grch38] [--pairs 4000000]       # index + FASTQ files (once per box)
  torchrun --nproc-per-node N tools/product_scale.py run [--preset ...] [--mode map|twopass]

`run` prints one JSON line on rank 0: wall seconds of the whole sharded run (FASTQ -> merged Aligned.out.sam / SJ.out.tab /
Log.final.out), pairs/s, and the per-phase times of rank 0 (star_b200.dist TIMING: mapping incl. index load, junction all-gather
bytes / seconds, 1st-pass junction merge, counter all-reduce + wait for the slowest rank, shard merge).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["prepare", "run"])
    ap.add_argument("--preset", default="chr21")
    ap.add_argument("--pairs", type=int, default=4_000_000)
    ap.add_argument("--mode", default="map", choices=["map", "twopass"])
    ap.add_argument("--threads", type=int, default=0, help="--runThreadN per rank (0: host cores / world size, at most 32)")
    ap.add_argument("--workdir", default=os.environ.get("STAR_B200_BENCH_DIR", "/tmp/star_b200_bench"))
    a = ap.parse_args()
    import bench
    import synth
    wd = os.path.join(a.workdir, a.preset)
    os.makedirs(wd, exist_ok=True)
    fq1, fq2 = os.path.join(wd, "prod_1.fq"), os.path.join(wd, "prod_2.fq")
    if a.cmd == "prepare":
        chrs, trs, idx, build = bench.prepare_genome(wd, a.preset)
        t0 = time.time()
        step = 1 << 20
        with open(fq1, "wb") as f1, open(fq2, "wb") as f2:   # written in blocks (names stay unique and ordered)
            for lo in range(0, a.pairs, step):
                n = min(step, a.pairs - lo)
                m1, m2 = synth.make_reads(chrs, trs, n, read_len=100, mm=0.005, seed=7000 + lo // step)
                t1, t2 = fq1 + ".part", fq2 + ".part"
                synth.write_fastq(m1, t1, first_index=lo)
                synth.write_fastq(m2, t2, first_index=lo)
                f1.write(open(t1, "rb").read())
                f2.write(open(t2, "rb").read())
        for t in (fq1 + ".part", fq2 + ".part"):
            if os.path.exists(t):
                os.remove(t)
        print(json.dumps({"prepared": a.preset, "pairs": a.pairs, "index_build": build, "fastq_seconds": round(time.time() - t0, 1)}), flush=True)
        return 0
    from star_b200 import dist as sd
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = a.threads or max(4, min(32, (os.cpu_count() or 8) // world))
    out = os.path.join(wd, "prod_%s_n%d/" % (a.mode, world))
    if rank == 0:
        import shutil
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
    pairs = sum(1 for _ in open(fq1, "rb")) // 4 if rank == 0 else 0
    argv = ["--genomeDir", os.path.join(wd, "idx"), "--readFilesIn", fq1, fq2, "--outFileNamePrefix", out, "--runThreadN", str(threads), "--outSAMtype", "SAM"]
    if a.mode == "twopass":
        argv += ["--twopassMode", "Basic"]
    t0 = time.time()
    rc = sd.run_sharded(argv)
    wall = time.time() - t0
    if rank == 0:
        timing = {}
        try:
            timing = json.load(open(out + "dist_timing.json"))
        except OSError:
            pass
        print(json.dumps({"product_path": "python -m star_b200.dist", "mode": a.mode, "preset": a.preset, "n_gpus": world, "pairs": pairs, "rc": rc,
                          "wall_s": round(wall, 2), "pairs_per_s_incl_startup": round(pairs / wall, 1), "threads_per_rank": threads, "rank0_phases": timing}), flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
