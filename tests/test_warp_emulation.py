"""The warp-uniform seed search of the next GPU kernel (star_b200/csrc/engine/seed_warp.cuh) compiled for the host: 32 threads play
the 32 lanes and meet at a barrier in every collective (oracle/warp_emul.cpp).  Every stored piece of every read must equal the
oracle's (= the reference's PC[] after seeding), i.e. the device logic is checked lane by lane without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT
EMUL_LIB = os.path.join(ROOT, "oracle", "_build", "libwarp_emul.so")


def _emul():
    if not os.path.exists(EMUL_LIB):
        oc.build_oracle()
    from star_b200 import capi
    lib = C.CDLL(EMUL_LIB)
    lib.warp_emul_seed_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("name,n_take,lmax", [("std", 250, 0), ("hard", 70, 0), ("se", 250, 0), ("hard", 40, 25)])
def test_emulated_warp_seed_search_equals_oracle_pieces(oracle, lib, golden, name, n_take, lmax):
    """lmax > 0: --seedSearchLmax, the extra fixed-length search from every start (ReadAlign_mapOneRead.cpp:81-87)."""
    import star_b200 as sb
    from star_b200 import capi
    files = [os.path.join(golden, name + "_1.fq")] + ([os.path.join(golden, name + "_2.fq")] if name != "se" else [])
    mates = [cf.read_fastq_seqs(f)[:n_take] for f in files]
    seq, off, n, nm = sb.pack_reads(mates)
    params = capi.default_params(lib)
    params.seedSearchLmax = lmax
    idx = sb.Index(lib, os.path.join(golden, "idx"), params=params)
    oe = oc.OracleEngine(oracle, idx)
    _, _, st_o, (pc_off_o, pc_o) = oe.map_chunk(seq, off, n, nm, dump=True)
    batch = oe._batch(seq, off, n, nm)
    oe.close()
    em = _emul()
    cap = int(pc_off_o[-1]) + 64 * n
    pc_off = np.zeros(n + 1, dtype=np.uint64)
    pc = np.zeros((cap, 8), dtype=np.uint64)
    per_read = np.zeros(4 * n, dtype=np.uint32)
    counters = np.zeros(4, dtype=np.uint64)
    rc = em.warp_emul_seed_chunk(idx.view, C.byref(idx.params), C.byref(batch), pc_off.ctypes.data, pc.ctypes.data, cap, per_read.ctypes.data, counters.ctypes.data)
    idx.close()
    assert rc == 0, "lanes disagreed on a uniform value (2) or capacity (1): %d" % rc
    assert np.array_equal(pc_off, pc_off_o)
    got = pc[: int(pc_off[-1])]
    # SAend of the oracle dump = SAstart + Nrep - 1; all eight fields must agree for every piece of every read
    bad = np.nonzero((got != pc_o).any(axis=1))[0]
    assert bad.size == 0, "first differing piece %d: emulated %s oracle %s" % (bad[0], got[bad[0]], pc_o[bad[0]])
    # the SAindex part of the search is unchanged, so these two counters equal the reference's
    assert int(counters[0]) == st_o.mmp_searches and int(counters[1]) == st_o.mmp_sai_words
    # and the point of the design: far fewer dependent rounds than the binary search's compare calls
    assert counters[3] * 2 < st_o.mmp_compare_calls


ENGINE_EMUL_LIB = os.path.join(ROOT, "oracle", "_build", "libengine_emul.so")


@pytest.mark.parametrize("name,n_take,lmax,env", [
    ("std", 400, 0, {}), ("hard", 120, 0, {}), ("se", 400, 0, {}), ("hard", 60, 25, {}), ("se", 300, 25, {}),
    ("std", 300, 0, {"STAR_B200_SEED_SCAN_MAX": "3"}),                                  # every window larger than 3 rows is bisected on the keys
    ("hard", 80, 25, {"STAR_B200_SEED_SCAN_MAX": "3", "STAR_B200_SEED_SORT_BITS": "0"}),  # ... chains in read order
    ("std", 200, 0, {"STAR_B200_SEED_RECS_PER_READ": "16"}),                            # record slabs overflow: those reads are flagged for the tier path
])
def test_emulated_keyed_seed_stage_equals_oracle_pieces(oracle, lib, golden, name, n_take, lmax, env, monkeypatch):
    """The DEFAULT seed stage (seed_keyed.cuh: SA keys, chain items sorted by SAindex L-mer, groups of 8 lanes, ordered replay) as emulated
    CTAs: every stored piece of every read equals the oracle's PC[] (= the reference's after seeding), and the algorithm-determined
    counters (searches, SAindex words) equal the instrumented oracle's."""
    import star_b200 as sb
    from star_b200 import capi
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    files = [os.path.join(golden, name + "_1.fq")] + ([os.path.join(golden, name + "_2.fq")] if name != "se" else [])
    mates = [cf.read_fastq_seqs(f)[:n_take] for f in files]
    seq, off, n, nm = sb.pack_reads(mates)
    params = capi.default_params(lib)
    params.seedSearchLmax = lmax
    idx = sb.Index(lib, os.path.join(golden, "idx"), params=params)
    oe = oc.OracleEngine(oracle, idx)
    _, _, st_o, (pc_off_o, pc_o) = oe.map_chunk(seq, off, n, nm, dump=True)
    batch = oe._batch(seq, off, n, nm)
    oe.close()
    if not os.path.exists(ENGINE_EMUL_LIB):
        oc.build_oracle()
    em = C.CDLL(ENGINE_EMUL_LIB)
    em.engine_emul_seed_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    cap = int(pc_off_o[-1]) + 64 * n
    pc_off = np.zeros(n + 1, dtype=np.uint64)
    pc = np.zeros((cap, 8), dtype=np.uint64)
    counters = np.zeros(4, dtype=np.uint64)
    rc = em.engine_emul_seed_chunk(idx.view, C.byref(idx.params), C.byref(batch), pc_off.ctypes.data, pc.ctypes.data, cap, counters.ctypes.data)
    idx.close()
    assert rc == 0
    if "STAR_B200_SEED_RECS_PER_READ" in env:   # flagged reads keep no pieces here (the tier path redoes them); the others must be exact
        assert 0 < int(counters[3]) < n
        cnt_e, cnt_o = np.diff(pc_off), np.diff(pc_off_o)
        flagged = cnt_e != cnt_o
        assert int(flagged.sum()) <= int(counters[3]) and (cnt_e[flagged] == 0).all()
        keep = np.repeat(~flagged, cnt_o.astype(np.int64))
        assert not (pc[: int(pc_off[-1])] != pc_o[keep]).any()
        return
    assert int(counters[3]) == 0
    assert np.array_equal(pc_off, pc_off_o)
    bad = np.nonzero((pc[: int(pc_off[-1])] != pc_o).any(axis=1))[0]
    assert bad.size == 0, "first differing piece %d: emulated %s oracle %s" % (bad[0], pc[bad[0]], pc_o[bad[0]])
    assert int(counters[0]) == st_o.mmp_searches and int(counters[1]) == st_o.mmp_sai_words


@pytest.mark.parametrize("name,n_take,env", [
    ("std", 24, {"STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "0"}),     # every read on the lane path (stitch_kernel)
    ("hard", 8, {"STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "0"}),
    ("se", 24, {"STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "0"}),
    ("std", 12, {}),                                                                       # DEFAULT pipeline: flat_setup -> flat_dfs_warp -> flat_record_warp
    ("hard", 1, {}),                                                                       # (2x150 bp at 5 % mismatches: thousands of sub-tree tasks per read)
    ("se", 12, {}),
    ("std", 8, {"STAR_B200_HEAVY_SPLIT": "2", "STAR_B200_FLAT_STORE_ALL": "0"}),           # many prefix sub-trees; leaves replayed by the recording kernel
    ("std", 8, {"ENGINE_EMUL_HOST_RECORD": "1"}),                                          # task kernel + sequential host restatement of the recording
    ("std", 12, {"STAR_B200_BIN_FILTER": "1"}),                                           # with the hashed (strand, bin) bitmap in front of the window lookup (optional: measured no gain)
    ("hard", 2, {"STAR_B200_BIN_FILTER": "1", "STAR_B200_SORTED_LOOKUP_MIN": "1"}),       # ... in front of the bisection
    ("std", 16, {"STAR_B200_SORTED_LOOKUP_MIN": "1"}),                                    # window of a locus by bisection over the sorted live windows (reads with many windows)
    ("hard", 4, {"STAR_B200_SORTED_LOOKUP_MIN": "1", "STAR_B200_HEAVY_FLAT": "0"}),        # ... in the warp-per-read kernel
    ("std", 12, {"STAR_B200_SEED_RECS_PER_READ": "16"}),                                  # reads flagged by the seed stage: re-seeded by the warp kernel, flat kernels again (overflow tier)
    ("hard", 3, {"STAR_B200_SEED_RECS_PER_READ": "16", "STAR_B200_FLAT_TIER": "0"}),       # ... the lane-per-read tier
    ("std", 16, {"STAR_B200_HEAVY_FLAT": "0"}),                                          # warp-per-read kernel, cooperative windows (mode B)
    ("hard", 10, {"STAR_B200_HEAVY_FLAT": "0", "STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "1"}),   # ... reads exported by their lane (mode A)
])
def test_emulated_kernels_equal_oracle(oracle, lib, golden, name, n_take, env, monkeypatch):
    """The UNMODIFIED kernel sources (seed.cu, stitch.cu, stitch_flat.cuh) compiled as host code through oracle/cuda_host_shim.h and run
    as emulated CTAs of host threads (warp collectives = per-warp barriers): every kernel of the default pipeline (prep_reads,
    seed_search, flat_setup, flat_dfs_warp, flat_record_warp), the lane path (stitch_kernel, overflow tier) and the warp-per-read
    kernel must give the oracle's alignments field by field.  Threads run freely between collectives, so this also checks that the
    warp-shared state of the flat kernels has a single writer and that every collective is reached by all 32 lanes."""
    import star_b200 as sb
    from star_b200 import capi
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if not os.path.exists(ENGINE_EMUL_LIB):
        oc.build_oracle()
    files = [os.path.join(golden, name + "_1.fq")] + ([os.path.join(golden, name + "_2.fq")] if name != "se" else [])
    mates = [cf.read_fastq_seqs(f)[:n_take] for f in files]
    seq, off, n, nm = sb.pack_reads(mates)
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    oe = oc.OracleEngine(oracle, idx)
    res_o, al_o, _ = oe.map_chunk(seq, off, n, nm)
    batch = oe._batch(seq, off, n, nm)
    res, al, ab = oe._out(n, oe.n_out)
    oe.close()
    em = C.CDLL(ENGINE_EMUL_LIB)
    em.engine_emul_map_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.c_void_p]
    info4 = np.zeros(4, dtype=np.uint64)
    rc = em.engine_emul_map_chunk(idx.view, C.byref(idx.params), C.byref(batch), C.byref(ab), info4.ctypes.data)
    idx.close()
    assert rc == 0 and int(info4[2]) == 0
    if "STAR_B200_HEAVY_NA" in env and "STAR_B200_HEAVY_FLAT" not in env:
        assert int(info4[0]) == 0 and int(info4[1]) == n
    elif "STAR_B200_FLAT_TIER" not in env:   # (with every read flagged by the seed stage, the tier is the only path taken)
        assert int(info4[0]) > 0, "no read reached the flat path / the warp-per-read kernel"
    diffs = oc.compare_outputs(res_o, al_o, res, al[:ab.nAligns])
    assert not diffs, "\n".join(diffs[:10])


@pytest.mark.parametrize("env", [{}, {"STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "0"}])
def test_emulated_kernels_sj_novel_filter(oracle, lib, golden, twopass_golden, env, monkeypatch):
    """2nd stage of --outFilterType BySJout (stitchWindowAligns.cpp:169-177) in the device code: with the list of surviving novel
    junctions set, evalLeaf drops transcripts whose unannotated junctions are not in it.  Emulated kernels (default pipeline and lane
    path) against the oracle on the un-annotated index, with every second junction of the unfiltered alignments in the list."""
    import star_b200 as sb
    from star_b200 import capi
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    all_mates = [cf.read_fastq_seqs(os.path.join(golden, "std_%d.fq" % m))[:300] for m in (1, 2)]
    idx = sb.Index(lib, os.path.join(twopass_golden, "idx0"))
    oe = oc.OracleEngine(oracle, idx)
    seq, off, n, nm = sb.pack_reads(all_mates)
    resA, alA, _ = oe.map_chunk(seq, off, n, nm)   # pick 6 spliced reads and 4 others (the emulation is slow)
    spliced = [i for i in range(n) if resA["nTrOut"][i] > 0 and any(
        (alA[int(resA["trOffset"][i]) + k]["canonSJ"][:max(0, int(alA[int(resA["trOffset"][i]) + k]["nExons"]) - 1)] >= 0).any() for k in range(int(resA["nTrOut"][i])))]
    pick = sorted(spliced[:6] + [i for i in range(n) if i not in spliced][:4])
    mates = [[m[i] for i in pick] for m in all_mates]
    seq, off, n, nm = sb.pack_reads(mates)
    res0, al0, _ = oe.map_chunk(seq, off, n, nm)
    sj = set()
    for a in al0:
        for iex in range(int(a["nExons"]) - 1):
            if a["canonSJ"][iex] >= 0 and a["sjAnnot"][iex] == 0:
                sj.add((int(a["exG"][iex]) + int(a["exL"][iex]), int(a["exG"][iex + 1]) - 1))
    sj = sorted(sj)
    assert len(sj) >= 3
    keep = sj[::2]
    s = np.array([x[0] for x in keep], dtype=np.uint64)
    e = np.array([x[1] for x in keep], dtype=np.uint64)
    oracle.star_oracle_set_sj_novel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    assert oracle.star_oracle_set_sj_novel(oe.ctx, s.ctypes.data, e.ctypes.data, len(keep)) == 0
    res_o, al_o, _ = oe.map_chunk(seq, off, n, nm)
    assert oc.compare_outputs(res0, al0, res_o, al_o), "the filter changed nothing: the test does not test"
    batch = oe._batch(seq, off, n, nm)
    res, al, ab = oe._out(n, oe.n_out)
    oe.close()
    em = C.CDLL(ENGINE_EMUL_LIB)
    em.engine_emul_map_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.c_void_p]
    em.engine_emul_set_sj_novel.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    em.engine_emul_set_sj_novel(s.ctypes.data, e.ctypes.data, len(keep))
    try:
        info4 = np.zeros(4, dtype=np.uint64)
        rc = em.engine_emul_map_chunk(idx.view, C.byref(idx.params), C.byref(batch), C.byref(ab), info4.ctypes.data)
    finally:
        em.engine_emul_set_sj_novel(None, None, C.c_uint64(2**64 - 1))
        idx.close()
    assert rc == 0 and int(info4[2]) == 0
    diffs = oc.compare_outputs(res_o, al_o, res, al[:ab.nAligns])
    assert not diffs, "\n".join(diffs[:10])


@pytest.mark.parametrize("env", [{}, {"STAR_B200_HEAVY_NA": "2000000000", "STAR_B200_HEAVY_EST": "0"}])
def test_emulated_kernels_take_empty_mates(oracle, lib, golden, env, monkeypatch):
    """A mate that was clipped to nothing before mapping (--clip3pNbases, adapters) reaches the engine as an empty mate of a pair: mate 1
    empty, mate 2 empty, both empty, a 19-base mate with an empty partner — emulated kernels equal the oracle on the flat and the lane path."""
    import star_b200 as sb
    from star_b200 import capi
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    m1 = cf.read_fastq_seqs(os.path.join(golden, "std_1.fq"))[:8]
    m2 = cf.read_fastq_seqs(os.path.join(golden, "std_2.fq"))[:8]
    parts, off = [], [0]
    for i, (a, b) in enumerate(zip(m1, m2)):
        a, b = bytes(a), bytes(b)
        if i % 3 == 0:
            b = b""
        if i % 4 == 1:
            a = b""
        if i == 5:
            a, b = a[:19], b""
        if i == 7:
            a, b = b"", b""
        for s in (a, b):
            parts.append(s)
            off.append(off[-1] + len(s))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    off = np.array(off, dtype=np.uint64)
    n, nm = 8, 2
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    oe = oc.OracleEngine(oracle, idx)
    res_o, al_o, _ = oe.map_chunk(seq, off, n, nm)
    batch = oe._batch(seq, off, n, nm)
    res, al, ab = oe._out(n, oe.n_out)
    oe.close()
    em = C.CDLL(ENGINE_EMUL_LIB)
    em.engine_emul_map_chunk.argtypes = [C.POINTER(capi.IndexView), C.POINTER(capi.Params), C.POINTER(capi.ReadBatch), C.POINTER(capi.AlignBatch), C.c_void_p]
    info4 = np.zeros(4, dtype=np.uint64)
    rc = em.engine_emul_map_chunk(idx.view, C.byref(idx.params), C.byref(batch), C.byref(ab), info4.ctypes.data)
    idx.close()
    assert rc == 0 and int(info4[2]) == 0
    assert (res_o["unmapType"] < 0).sum() >= 4
    diffs = oc.compare_outputs(res_o, al_o, res, al[:ab.nAligns])
    assert not diffs, "\n".join(diffs[:10])
