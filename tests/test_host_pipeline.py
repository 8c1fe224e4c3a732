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


def _variant(lines, kind):
    out = list(lines)
    n = len(out) // 4
    if kind == "crlf":
        return "\r\n".join(out[:4 * n]) + "\r\n"
    if kind == "no_final_newline":
        return "\n".join(out[:4 * n])
    if kind == "trailing_blank_lines":
        return "\n".join(out[:4 * n]) + "\n\n\n"
    if kind == "comments_and_filter_flags":
        for r in range(n):
            out[4 * r] = out[4 * r].split()[0] + (" 1:Y:0:ACGT" if r % 3 == 0 else " 2:N:18:ACGT extra words")
            if r % 5 == 0:
                out[4 * r + 1] = out[4 * r + 1].lower()
        return "\n".join(out[:4 * n]) + "\n"
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["crlf", "no_final_newline", "trailing_blank_lines", "comments_and_filter_flags"])
def test_mapped_file_reader_equals_the_stream_reader(oracle, golden, tmp_path, kind):
    """Plain FASTQ files go through the memory-mapped parallel parser, piped input (--readFilesCommand) through the line-by-line stream
    parser (the restatement of processChunks/readLoad): same records, names, filter flags, order — on awkward but legal text."""
    files = []
    for m in (1, 2):
        with open(os.path.join(golden, "std_%d.fq" % m)) as f:
            lines = f.read().split("\n")
        while lines and lines[-1] == "":
            lines.pop()
        p = str(tmp_path / ("v_%d.fq" % m))
        with open(p, "w", newline="") as f:
            f.write(_variant(lines[:4 * 600], kind))
        files.append(p)
    extra = ["--gpuChunkReads", "97", "--outSAMunmapped", "Within", "--readNameSeparator", "/", "_"]
    a, b = str(tmp_path) + "/fast.", str(tmp_path) + "/stream."
    _cli(os.path.join(golden, "idx"), files, a, extra, threads=4)
    _cli(os.path.join(golden, "idx"), files, b, extra + ["--readFilesCommand", "cat"], threads=4)
    assert "reads input" in open(a + "Log.out").read()
    sa, sb_ = cf.sam_body(a + "Aligned.out.sam"), cf.sam_body(b + "Aligned.out.sam")
    assert len(sa) > 1000 and sa == sb_
    assert open(a + "SJ.out.tab", "rb").read() == open(b + "SJ.out.tab", "rb").read()
    assert cf.log_counters(a + "Log.final.out") == cf.log_counters(b + "Log.final.out")
    if kind == "comments_and_filter_flags":
        flags = [int(l.split(b"\t")[1]) for l in sa]
        assert any(f & 0x200 for f in flags) and not all(f & 0x200 for f in flags)
