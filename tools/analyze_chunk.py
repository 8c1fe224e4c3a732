#!/usr/bin/env python3
"""Per-read work analysis of one chunk on the GPU (uses star_gpu_debug_read_info). Usage: analyze_chunk.py [pairs] [mm] [readlen]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench
import star_b200 as sb
import synth

INFO = np.dtype({"names": ["Lread", "rl0", "rl1", "nP", "nA", "mapMarker", "multNminL", "Nsplit", "split1_0", "mmTotal", "flags",
                           "searches", "saiWords", "compare", "bases", "saEnum", "nodes", "leaves", "slow"],
                 "formats": ["<u4", "<u2", "<u2", "<u2", "<u4", "<u4", "<u4", "<u2", "<u2", "<u4", "<u4"] + ["<u4"] * 8,
                 "offsets": [0, 4, 6, 8, 12, 16, 20, 24, 26, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64], "itemsize": 68})

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
mm = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
rl = int(sys.argv[3]) if len(sys.argv) > 3 else 100
preset = os.environ.get("STAR_B200_BENCH_PRESET", "chr21")
workdir = os.path.join(os.environ.get("STAR_B200_BENCH_DIR", "/tmp/star_b200_bench"), preset)
os.makedirs(workdir, exist_ok=True)
chrs, trs, idx, _ = bench.prepare_genome(workdir, preset)
lib = sb.load_library()
lib.star_gpu_debug_read_info.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
index = sb.Index(lib, idx)
m1, m2 = synth.make_reads(chrs, trs, n, read_len=rl, mm=mm, seed=1000)
seq, off, _, nm = sb.pack_reads([m1, m2])
eng = sb.Engine(lib, index, max_reads=n)
eng.upload(seq, off, n, nm)
lib.star_gpu_debug_prof.argtypes = [C.c_void_p, C.c_void_p]
for k in range(3):
    st = eng.map_resident()
    prof = np.zeros(32, dtype=np.uint64)
    lib.star_gpu_debug_prof(eng.ctx, prof.ctypes.data)
    tot = float(prof[:8].sum()) or 1.0
    print("light kernel warp-cycles by phase [fetch+copy, win, flank, assign, nextwin, node, leaf, select]: " + " ".join("%.1f%%" % (100.0 * float(x) / tot) for x in prof[:8]) + "  total %.3g" % tot)
    th = float(prof[16:19].sum()) or 1.0
    # flat path (default): slots 16/17/18 = summed warp-cycles of flat_setup_kernel / flat_dfs_warp_kernel / flat_record_warp_kernel (the kernels run
    # different numbers of warps: divide by the warp count of each grid for times); warp-per-read kernel (STAR_B200_HEAVY_FLAT=0): its setup / E / R phases
    print("heavy kernel warp-cycles [setup, E, R]: " + " ".join("%.1f%%" % (100.0 * float(x) / th) for x in prof[16:19]) + "  total %.3g; tasks %d replays %d" % (th, prof[19], prof[20]))
    if prof[21]:
        it = float(prof[21])
        print("warp-per-read kernel E-phase: iterations/warp-sum %.3g, avg active lanes: fetch %.2f node %.2f leaf %.2f" % (it, prof[22] / it, prof[23] / it, prof[24] / it))
    print("run", k, {k2: round(v, 2) if isinstance(v, float) else v for k2, v in st.as_dict().items()})
info = np.zeros(n, dtype=INFO)
rc = lib.star_gpu_debug_read_info(eng.ctx, info.ctypes.data, info.nbytes)
assert rc == 0
slow = info["slow"] > 0
print("pairs/s total %.0f ; fast-only stitch %.0f ; seed %.0f" % (n / st.ms_total * 1e3, n / max(st.ms_stitch, 1e-3) * 1e3, n / st.ms_seed * 1e3))
for name in ["nP", "saEnum", "nodes", "leaves", "compare"]:
    x = info[name].astype(np.float64)
    q = np.percentile(x, [50, 90, 99, 99.9, 100])
    print("%8s mean %.1f  p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f | slow-path reads mean %.1f (n=%d) | share of total work in slow reads %.3f"
          % (name, x.mean(), q[0], q[1], q[2], q[3], q[4], x[slow].mean() if slow.any() else 0, slow.sum(), x[slow].sum() / max(1, x.sum())))
eng.close() if not os.environ.get("ANALYZE_LIGHT") else eng.close()

# ---- throughput of the light reads alone (tail excluded): keeps only reads below the p-th percentile of nA
if os.environ.get("ANALYZE_LIGHT"):
    for pct in (90, 98, 99.5):
        thr = np.percentile(info["nA"], pct)
        keep = np.nonzero(info["nA"] <= thr)[0]
        m1k, m2k = m1[keep], m2[keep]
        seq2, off2, n2, _ = sb.pack_reads([m1k, m2k])
        eng = sb.Engine(lib, index, max_reads=n2)
        eng.upload(seq2, off2, n2, nm)
        for k in range(2):
            st = eng.map_resident()
        print("light reads <= p%.1f of nA (nA<=%d): n=%d  seed %.1f ms  fast stitch %.1f ms  tiers %.1f ms -> %.0f pairs/s (fast stitch only %.0f); nodes %d slow %d"
              % (pct, thr, n2, st.ms_seed, st.ms_stitch, st.ms_window, n2 / st.ms_total * 1e3, n2 / st.ms_stitch * 1e3, st.stitch_nodes, st.slow_path_reads))
        eng.close()
rs = (info["flags"] >> 8) & 0xff
print("overflow reasons (1 windows, 3 transcripts, 4 pool, 5 heavy scratch):", {int(r): int((rs == r).sum()) for r in np.unique(rs) if r})
