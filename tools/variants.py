#!/usr/bin/env python3
"""One process, one index load: engine contexts created under different STAR_B200_* settings, 1 M pairs mapped resident, per-stage times.
usage: python tools/variants.py [preset] [pairs] -- NAME=VAL[,NAME=VAL...] ...   (each argument after -- is one variant; 'base' = defaults)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
import star_b200 as sb
import synth

args = sys.argv[1:]
split = args.index("--") if "--" in args else len(args)
preset = args[0] if split > 0 else "grch38"
n = int(args[1]) if split > 1 else 1 << 20
variants = args[split + 1:] or ["base"]
wd = os.path.join(os.environ.get("STAR_B200_BENCH_DIR", "/tmp/star_b200_bench"), preset)
chrs, trs, idx, _ = bench.prepare_genome(wd, preset)
lib = sb.load_library()
index = sb.Index(lib, idx)
m1, m2 = synth.make_reads(chrs, trs, n, read_len=100, mm=0.005, seed=1000)
del chrs
seq, off, _, nm = sb.pack_reads([m1, m2])
for v in variants:
    sets = {} if v == "base" else dict(kv.split("=") for kv in v.split(","))
    for k, val in sets.items():
        os.environ[k] = val
    eng = sb.Engine(lib, index, max_reads=n)
    eng.upload(seq, off, n, nm)
    best = None
    for _ in range(4):
        st = eng.map_resident()
        if best is None or st.ms_total < best["ms_total"]:
            best = {"ms_total": st.ms_total, "ms_seed": st.ms_seed, "ms_stitch": st.ms_stitch, "ms_tiers": st.ms_window, "ms_pack": st.ms_pack}
    eng.close()
    for k in sets:
        del os.environ[k]
    print(json.dumps({"variant": v, **{k: round(x, 2) for k, x in best.items()}}), flush=True)
