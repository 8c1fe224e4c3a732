#!/usr/bin/env python3
"""Runs the emulated kernels (oracle/engine_emul.cpp: the unmodified kernel sources as CTAs of host threads) from a sanitizer build.

  make -f oracle/Makefile asan
  LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS=halt_on_error=0 [STAR_B200_...=...] \\
      python tools/sanitize_emul.py <unpacked tests/golden/tiny.tar.gz>/tiny std 6 oracle/_build/asan/libengine_emul_tsan.so
  (libasan.so / libengine_emul_asan.so for AddressSanitizer).  Sets: std | hard | se; the number is how many reads are mapped.
Lanes are host threads that run freely between collectives, so ThreadSanitizer checks the single-writer discipline of the warp-shared state and that
every collective is reached by all lanes of its mask."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as cf
import star_b200 as sb
from star_b200 import capi

golden, name, n_take, libpath = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
lib = sb.load_library()
files = [os.path.join(golden, name + "_1.fq")] + ([os.path.join(golden, name + "_2.fq")] if name != "se" else [])
mates = [cf.read_fastq_seqs(f)[:n_take] for f in files]
seq, off, n, nm = sb.pack_reads(mates)
idx = sb.Index(lib, os.path.join(golden, "idx"))


class _E(capi._EngineBase):
    pass


e = _E()
batch = e._batch(seq, off, n, nm)
res, al, ab = e._out(n, max(1, int(idx.params.outFilterMultimapNmax)))
em = C.CDLL(libpath)
em.engine_emul_map_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.c_void_p]
info4 = np.zeros(4, dtype=np.uint64)
rc = em.engine_emul_map_chunk(idx.view, C.byref(idx.params), C.byref(batch), C.byref(ab), info4.ctypes.data)
print("rc", rc, "alignments", ab.nAligns, "reads on [warp/flat path, lane path, failed, launches]:", info4)
sys.exit(1 if rc else 0)
