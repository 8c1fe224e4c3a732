"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(star_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from star_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
ORACLE_CLI = os.path.join(ROOT, "oracle", "_build", "star_cli_oracle")
REF_STAR = os.path.join(ROOT, "oracle", "_ref", "STAR")


class Dump(C.Structure):
    _fields_ = [("pcOff", C.POINTER(C.c_uint64)), ("pc", C.POINTER(C.c_uint64))]


def build_oracle():
    subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile"), "-j8"], cwd=ROOT)


def load_oracle():
    if not os.path.exists(ORACLE_LIB):
        build_oracle()
    lib = C.CDLL(ORACLE_LIB)
    lib.star_oracle_init.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.c_uint32]
    lib.star_oracle_map_chunk.argtypes = [C.c_void_p, C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.POINTER(capi.ChunkStats)]
    lib.star_oracle_map_chunk_dump.argtypes = [C.c_void_p, C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.POINTER(capi.ChunkStats), C.POINTER(Dump)]
    lib.star_oracle_dump_free.argtypes = [C.POINTER(Dump)]
    lib.star_oracle_destroy.argtypes = [C.c_void_p]
    lib.star_oracle_last_error.restype = C.c_char_p
    return lib


class OracleEngine(capi._EngineBase):
    def __init__(self, olib, index):
        self.olib = olib
        self.index = index
        self.n_out = max(1, int(index.params.outFilterMultimapNmax))
        ctx = C.c_void_p()
        rc = olib.star_oracle_init(C.byref(ctx), 0, index.view, C.byref(index.params), 0)
        assert rc == 0
        self.ctx = ctx

    def map_chunk(self, seq, off, n, n_mates, dump=False):
        b = self._batch(seq, off, n, n_mates)
        res, al, ab = self._out(n, self.n_out)
        st = capi.ChunkStats()
        if dump:
            d = Dump()
            rc = self.olib.star_oracle_map_chunk_dump(self.ctx, C.byref(b), C.byref(ab), C.byref(st), C.byref(d))
        else:
            rc = self.olib.star_oracle_map_chunk(self.ctx, C.byref(b), C.byref(ab), C.byref(st))
        if rc:
            raise capi.StarError(rc, self.olib.star_oracle_last_error().decode())
        if dump:
            pc_off = np.ctypeslib.as_array(d.pcOff, shape=(n + 1,)).copy()
            pc = np.ctypeslib.as_array(d.pc, shape=(max(1, int(pc_off[-1])) * 8,)).copy()[: int(pc_off[-1]) * 8].reshape(-1, 8)
            self.olib.star_oracle_dump_free(C.byref(d))
            return res, al[:ab.nAligns], st, (pc_off, pc)
        return res, al[:ab.nAligns], st

    def close(self):
        if self.ctx:
            self.olib.star_oracle_destroy(self.ctx)
            self.ctx = None


def compare_outputs(res_a, al_a, res_b, al_b):
    """Field-wise bit-exact comparison; returns a list of human-readable differences (empty = identical)."""
    diffs = []
    if len(res_a) != len(res_b):
        return ["different number of reads: %d vs %d" % (len(res_a), len(res_b))]
    for name in res_a.dtype.names:
        bad = np.nonzero(res_a[name] != res_b[name])[0]
        if len(bad):
            diffs.append("result.%s differs for %d reads, first read %d: %s vs %s" % (name, len(bad), bad[0], res_a[name][bad[0]], res_b[name][bad[0]]))
    if len(al_a) != len(al_b):
        diffs.append("different number of alignments: %d vs %d" % (len(al_a), len(al_b)))
        return diffs
    if len(al_a) == 0:   # (no read of the chunk mapped: nothing to compare field by field)
        return diffs
    for name in al_a.dtype.names:
        x, y = al_a[name], al_b[name]
        neq = x != y
        if neq.ndim > 1:
            neq = neq.reshape(len(x), -1).any(axis=1)
        bad = np.nonzero(neq)[0]
        if len(bad):
            diffs.append("align.%s differs for %d alignments, first #%d: %s vs %s" % (name, len(bad), bad[0], x[bad[0]], y[bad[0]]))
    return diffs
