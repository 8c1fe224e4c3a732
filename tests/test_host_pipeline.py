"""CPU tests of the host side around the engine (the drop-in command line driven by the oracle engine): the three-stage chunk
pipeline, input errors in a late chunk, read sharding and the shard merge.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT


def _cli(genome, files, out, extra=(), threads=3, check=True):
    cmd = [oc.ORACLE_CLI, "--genomeDir", genome, "--readFilesIn"] + files + ["--outFileNamePrefix", out, "--runThreadN", str(threads)] + list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=check)


@pytest.mark.parametrize("chunk", [1, 7, 256, 5000])
def test_outputs_do_not_depend_on_the_chunk_size(oracle, golden, tmp_path, chunk):
    """Chunks flow through reader / engine / output threads with 3 buffers in flight: order and content must not change."""
    out = str(tmp_path) + "/"
    n = 300 if chunk < 256 else 0   # tiny chunks: a prefix of the reads is enough (and keeps the test short)
    extra = ["--gpuChunkReads", str(chunk)] + (["--readMapNumber", str(n)] if n else [])
    _cli(os.path.join(golden, "idx"), [os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq")], out, extra)
    ref = cf.sam_body(os.path.join(golden, "ref_std", "Aligned.out.sam"))
    ours = cf.sam_body(out + "Aligned.out.sam")
    if n:   # the first n reads = the first n record names of the FASTQ
        with open(os.path.join(golden, "std_1.fq")) as f:
            lines = f.read().split("\n")
        keep = set(lines[4 * i][1:].split()[0].split("/")[0].encode() for i in range(n))
        ref = [l for l in ref if l.split(b"\t")[0] in keep]
    assert ours == ref


def test_input_error_in_a_late_chunk_stops_the_run(oracle, golden, tmp_path):
    """readLoad.cpp:66-71: quality length != sequence length is fatal; with the pipeline the error surfaces from the reader thread
    while earlier chunks are still being mapped / written — the process must exit with the reference's code, not hang."""
    with open(os.path.join(golden, "se_1.fq")) as f:
        lines = f.read().split("\n")
    bad = list(lines)
    rec = 150
    bad[4 * rec + 3] = bad[4 * rec + 3][:-3]           # truncate one quality string
    fq = str(tmp_path / "bad.fq")
    open(fq, "w").write("\n".join(bad))
    r = _cli(os.path.join(golden, "idx"), [fq], str(tmp_path) + "/o/", ["--gpuChunkReads", "64"], check=False)
    assert r.returncode != 0
    assert "quality string length is not equal to sequence length" in r.stderr
    assert "FATAL ERROR, exiting" in r.stderr


@pytest.mark.parametrize("world", [3])
def test_shards_partition_the_reads_in_order(oracle, golden, tmp_path, world):
    """--gpuShardIndex/--gpuShardCount: contiguous slices by record index, every read in exactly one shard, global read numbering kept."""
    bodies = []
    for r in range(world):
        out = str(tmp_path) + "/s%d." % r
        _cli(os.path.join(golden, "idx"), [os.path.join(golden, "se_1.fq")], out, ["--gpuShardIndex", str(r), "--gpuShardCount", str(world), "--outSAMreadID", "Number"])
        assert os.path.exists(out + "shard.bin") and not os.path.exists(out + "SJ.out.tab")
        sam = open(out + "Aligned.out.sam").read().split("\n")
        assert (sam[0].startswith("@")) == (r == 0)     # only shard 0 carries the header
        bodies.append(cf.sam_body(out + "Aligned.out.sam"))
    whole = str(tmp_path) + "/w."
    _cli(os.path.join(golden, "idx"), [os.path.join(golden, "se_1.fq")], whole, ["--outSAMreadID", "Number"])
    assert sum(bodies, []) == cf.sam_body(whole + "Aligned.out.sam")


def test_merge_without_allreduced_counters_sums_the_shard_files(oracle, lib, golden, tmp_path):
    """star_host_merge_shards(counters24 = NULL): the counters come from the shard files; outputs equal the single-process reference."""
    world = 2
    pre = str(tmp_path) + "/m_"
    args = ["--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq")]
    for r in range(world):
        subprocess.check_call([oc.ORACLE_CLI] + args + ["--outFileNamePrefix", pre + "shard%d." % r, "--gpuShardIndex", str(r), "--gpuShardCount", str(world)],
                              stdout=subprocess.DEVNULL, timeout=300)
    argv = ["STAR"] + args + ["--outFileNamePrefix", pre]
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    lib.star_host_merge_shards.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_void_p]
    assert lib.star_host_merge_shards(len(argv), arr, world, None) == 0
    ref = os.path.join(golden, "ref_std")
    assert cf.sam_body(pre + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert open(pre + "SJ.out.tab", "rb").read() == open(os.path.join(ref, "SJ.out.tab"), "rb").read()
    assert cf.log_counters(pre + "Log.final.out") == cf.log_counters(os.path.join(ref, "Log.final.out"))
