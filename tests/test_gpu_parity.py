"""GPU parity tests: the CUDA engine, called through the C-ABI, against the CPU oracle and the reference goldens.

Bar: bit-exact (integer/index work).  Every field of every returned alignment and read result must be equal.
"""
import os
import subprocess

import numpy as np
import pytest

import conftest as cf

pytestmark = pytest.mark.gpu

ROOT = cf.ROOT


def _sets(golden):
    return {
        "std": [os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq")],
        "hard": [os.path.join(golden, "hard_1.fq"), os.path.join(golden, "hard_2.fq")],
        "se": [os.path.join(golden, "se_1.fq")],
    }


@pytest.fixture(scope="module")
def tiny_index(lib, golden):
    import star_b200 as sb
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    yield idx
    idx.close()


@pytest.mark.parametrize("name", ["std", "hard", "se"])
def test_engine_matches_oracle_tiny(lib, oracle, golden, tiny_index, name):
    import oracle_capi as oc
    import star_b200 as sb
    files = _sets(golden)[name]
    mates = [cf.read_fastq_seqs(f) for f in files]
    seq, off, n, nm = sb.pack_reads(mates)
    oe = oc.OracleEngine(oracle, tiny_index)
    res_o, al_o, st_o = oe.map_chunk(seq, off, n, nm)
    oe.close()
    eng = sb.Engine(lib, tiny_index, max_reads=n)
    res_g, al_g, st_g = eng.map_chunk(seq, off, n, nm)
    eng.close()
    diffs = oc.compare_outputs(res_o, al_o, res_g, al_g)
    assert not diffs, "\n".join(diffs[:20])
    # the algorithm-determined work counters of the MMP search must agree with the instrumented oracle: searches and
    # SAindex words.  (Compare calls / bases examined belong to the reference's binary search; the keyed kernel probes differently and
    # reports its own probes there — the roofline numerator always takes the ORACLE's counts.)
    for k in ("mmp_searches", "mmp_sai_words", "sa_enumerated"):
        assert getattr(st_g, k) == getattr(st_o, k), k


@pytest.mark.parametrize("env", [
    {"STAR_B200_HEAVY_EST": "0"},                                   # heavy path off: everything on the one-lane-per-read path
    {"STAR_B200_HEAVY_NA": "0x7fffffff", "STAR_B200_HEAVY_EST": "1"},   # every read exported by its lane (mode A) -> flattened heavy path
    {"STAR_B200_HEAVY_NA": "1"},                                    # every read with >1 locus: cooperative windows (mode B) -> flattened heavy path
    {"STAR_B200_HEAVY_NA": "1", "STAR_B200_HEAVY_SPLIT": "2"},      # the same with windows cut into many prefix sub-trees
    {"STAR_B200_HEAVY_NA": "1", "STAR_B200_FLAT_MAXTASKS": "3000", "STAR_B200_FLAT_MAXBLOCKS": "40", "STAR_B200_FLAT_TRWORDS": "4096",
     "STAR_B200_FLAT_POOL_BYTES": "2000000"},                       # flat pools exhausted: reads fall to the overflow tiers / replay their leaves
    {"STAR_B200_HEAVY_FLAT": "0", "STAR_B200_HEAVY_NA": "0x7fffffff", "STAR_B200_HEAVY_EST": "1"},   # warp-per-read kernel, mode A
    {"STAR_B200_HEAVY_FLAT": "0", "STAR_B200_HEAVY_NA": "1"},       # warp-per-read kernel, mode B
    {"STAR_B200_HEAVY_FLAT": "0", "STAR_B200_HEAVY_NA": "1", "STAR_B200_HEAVY_MEMO": "256"},     # the same with the (optional) shared stitch memo switched on
    {"STAR_B200_FAST_MAXW": "4", "STAR_B200_FAST_MAXTR": "4", "STAR_B200_FAST_MAXP": "8", "STAR_B200_MID_MAXW": "16", "STAR_B200_MID_MAXTR": "8"},  # tiny caps: overflow tiers
])
@pytest.mark.parametrize("name", ["std", "hard"])
def test_engine_paths_are_all_exact(lib, oracle, golden, tiny_index, name, env):
    """Every execution path of the engine (light lanes, flattened heavy path, warp-per-read kernel, overflow tiers) must give the oracle's result."""
    import oracle_capi as oc
    import star_b200 as sb
    files = _sets(golden)[name]
    mates = [cf.read_fastq_seqs(f) for f in files]
    seq, off, n, nm = sb.pack_reads(mates)
    oe = oc.OracleEngine(oracle, tiny_index)
    res_o, al_o, st_o = oe.map_chunk(seq, off, n, nm)
    oe.close()
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = str(int(v, 0))
        eng = sb.Engine(lib, tiny_index, max_reads=n)
        res_g, al_g, st_g = eng.map_chunk(seq, off, n, nm)
        eng.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    diffs = oc.compare_outputs(res_o, al_o, res_g, al_g)
    assert not diffs, "\n".join(diffs[:20])
    for k in ("mmp_searches", "mmp_sai_words", "sa_enumerated"):
        assert getattr(st_g, k) == getattr(st_o, k), k


@pytest.mark.parametrize("name,extra", [("std", []), ("hard", []), ("se", []), ("std_opts", None)])
def test_cli_matches_reference_golden(lib, golden, tmp_path, name, extra):
    """Drop-in CLI on the GPU vs outputs of the unmodified reference binary (committed goldens)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    if extra is None:
        mg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mg)
        extra = mg.OPTS
    base = "std" if name == "std_opts" else name
    files = _sets(golden)[base]
    out = str(tmp_path) + "/"
    cmd = [os.path.join(ROOT, "star_b200", "bin", "STAR"), "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn"] + files + \
          ["--outFileNamePrefix", out, "--runThreadN", "2"] + list(extra)
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    ref = os.path.join(golden, "ref_" + name)
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(out + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(out + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))


@pytest.mark.parametrize("base,extra", [
    ("std", ["--outSAMattributes", "NH", "HI", "AS", "nM", "XS"]),
    ("std", ["--outSAMstrandField", "intronMotif", "--outFilterIntronMotifs", "RemoveNoncanonical"]),
    ("std", ["--alignEndsType", "Extend5pOfRead1", "--outSAMprimaryFlag", "AllBestScore"]),
    ("std", ["--outFilterMultimapNmax", "3", "--winAnchorMultimapNmax", "100", "--outSAMmultNmax", "2"]),
    ("hard", ["--outFilterMismatchNoverLmax", "0.1", "--scoreGenomicLengthLog2scale", "0", "--alignSJoverhangMin", "8"]),
])
def test_cli_option_sets_match_oracle_cli(lib, oracle, golden, tmp_path, base, extra):
    """Option sets without committed goldens: the drop-in CLI on the GPU vs the same host code driven by the oracle engine
    (the oracle itself is pinned to the unmodified reference for these option sets by tests/test_oracle_golden.py)."""
    import oracle_capi as oc
    files = _sets(golden)[base]
    outs = []
    for tag, binary in (("gpu", os.path.join(ROOT, "star_b200", "bin", "STAR")), ("ora", oc.ORACLE_CLI)):
        out = str(tmp_path / tag) + "/"
        os.makedirs(out)
        cmd = [binary, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn"] + files + ["--outFileNamePrefix", out, "--runThreadN", "2"] + list(extra)
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        outs.append(out)
    assert cf.sam_body(outs[0] + "Aligned.out.sam") == cf.sam_body(outs[1] + "Aligned.out.sam")
    assert open(outs[0] + "SJ.out.tab", "rb").read() == open(outs[1] + "SJ.out.tab", "rb").read()
    assert cf.log_counters(outs[0] + "Log.final.out") == cf.log_counters(outs[1] + "Log.final.out")


def test_cli_bam_outputs_match_oracle_cli(lib, oracle, golden, tmp_path):
    """--outSAMtype BAM Unsorted SortedByCoordinate on the GPU: decompressed records equal those of the same host code driven by the
    oracle engine (which tests/test_bam_output.py pins to the unmodified reference record for record)."""
    import gzip
    import oracle_capi as oc
    files = _sets(golden)["hard"]
    recs = {}
    for tag, binary in (("gpu", os.path.join(ROOT, "star_b200", "bin", "STAR")), ("ora", oc.ORACLE_CLI)):
        out = str(tmp_path / tag) + "/"
        os.makedirs(out)
        cmd = [binary, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn"] + files + ["--outFileNamePrefix", out, "--runThreadN", "3",
               "--outSAMtype", "BAM", "Unsorted", "SortedByCoordinate", "--outSAMunmapped", "Within", "--outSAMattributes", "All"]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        got = []
        for name in ("Aligned.out.bam", "Aligned.sortedByCoord.out.bam"):
            d = gzip.decompress(open(out + name, "rb").read())
            lt = int.from_bytes(d[4:8], "little")
            o = 8 + lt
            nref = int.from_bytes(d[o:o + 4], "little")
            o += 4
            for _ in range(nref):
                ln = int.from_bytes(d[o:o + 4], "little")
                o += 8 + ln
            got.append(d[o:])
        recs[tag] = got
    assert len(recs["gpu"][0]) > 100000
    assert recs["gpu"][0] == recs["ora"][0]
    assert recs["gpu"][1] == recs["ora"][1]


def test_init_failure_is_a_clean_error_and_leaves_nothing_behind(lib, oracle, golden, tiny_index):
    """A context whose pools cannot fit in HBM: star_gpu_init returns STAR_EXIT_RUNTIME with a CUDA message, frees what it had allocated
    (free memory before == after), and the next context on the same device works and is still exact."""
    import torch
    import oracle_capi as oc
    import star_b200 as sb
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    with pytest.raises(sb.StarError) as e:
        sb.Engine(lib, tiny_index, max_reads=400_000_000)      # hundreds of GB of per-read slabs
    assert e.value.code == 103 and "CUDA error" in str(e.value)
    free1, _ = torch.cuda.mem_get_info(0)
    assert free1 >= free0 - (64 << 20), "device memory of the failed init was not released: %d MB" % ((free0 - free1) >> 20)
    mates = [cf.read_fastq_seqs(f)[:200] for f in _sets(golden)["std"]]
    seq, off, n, nm = sb.pack_reads(mates)
    for _ in range(2):                                             # two init / destroy cycles: nothing accumulates
        eng = sb.Engine(lib, tiny_index, max_reads=n)
        res_g, al_g, _ = eng.map_chunk(seq, off, n, nm)
        eng.close()
    free2, _ = torch.cuda.mem_get_info(0)
    assert free2 >= free0 - (64 << 20)
    oe = oc.OracleEngine(oracle, tiny_index)
    res_o, al_o, _ = oe.map_chunk(seq, off, n, nm)
    oe.close()
    assert not oc.compare_outputs(res_o, al_o, res_g, al_g)


def test_page_locked_buffers_and_second_fetch(lib, golden, tiny_index):
    """star_gpu_host_alloc / star_gpu_host_free hand out page-locked memory the chunk calls accept; a record buffer that is too small makes
    star_gpu_map_chunk return STAR_EXIT_RUNTIME with out.nAligns = the capacity needed and keeps the results resident, and
    star_gpu_download_results into a large enough buffer returns what a call with a large buffer returns."""
    import ctypes as C
    import star_b200 as sb
    from star_b200 import capi
    lib.star_gpu_host_alloc.restype = C.c_void_p
    lib.star_gpu_host_alloc.argtypes = [C.c_size_t]
    lib.star_gpu_host_free.argtypes = [C.c_void_p]
    mates = [cf.read_fastq_seqs(f) for f in _sets(golden)["std"]]
    seq, off, n, nm = sb.pack_reads(mates)
    eng = sb.Engine(lib, tiny_index, max_reads=n)
    res0, al0, _ = eng.map_chunk(seq, off, n, nm)
    assert al0.shape[0] > 8
    # sequences, offsets, results and records in page-locked memory
    sizes = [seq.nbytes, off.nbytes, n * capi.RESULT_DTYPE.itemsize, al0.shape[0] * capi.ALIGN_DTYPE.itemsize]
    ptrs = [lib.star_gpu_host_alloc(s) for s in sizes]
    assert all(ptrs)
    try:
        C.memmove(ptrs[0], seq.ctypes.data, seq.nbytes)
        C.memmove(ptrs[1], off.ctypes.data, off.nbytes)
        b = capi.ReadBatch()
        b.nReads, b.nMates, b.seq, b.seqOff = n, nm, ptrs[0], ptrs[1]
        ab = capi.AlignBatch()
        ab.reads, ab.aligns, ab.alignsCapacity, ab.nAligns = ptrs[2], ptrs[3], 8, 0      # too small on purpose
        st = capi.ChunkStats()
        rc = lib.star_gpu_map_chunk(eng.ctx, C.byref(b), C.byref(ab), C.byref(st))
        assert rc != 0 and ab.nAligns == al0.shape[0]
        ab.alignsCapacity, ab.nAligns = al0.shape[0], 0
        assert lib.star_gpu_download_results(eng.ctx, C.byref(ab)) == 0 and ab.nAligns == al0.shape[0]
        res1 = np.frombuffer((C.c_char * sizes[2]).from_address(ptrs[2]), dtype=capi.RESULT_DTYPE).copy()
        al1 = np.frombuffer((C.c_char * sizes[3]).from_address(ptrs[3]), dtype=capi.ALIGN_DTYPE).copy()
        import oracle_capi as oc
        diffs = oc.compare_outputs(res0, al0, res1, al1)     # field by field (padding bytes are not part of the contract)
        assert not diffs, "\n".join(diffs[:10])
    finally:
        for p in ptrs:
            lib.star_gpu_host_free(p)
        eng.close()
