"""--outSAMtype BAM Unsorted: the decompressed BAM stream of the host code (driven by the oracle engine on CPU)
must equal the unmodified reference's record for record; BGZF framing is checked structurally (block sizes, EOF marker).  No GPU."""
import ctypes as C
import gzip
import os
import struct
import subprocess

import pytest

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT


def parse_bam(path):
    raw = open(path, "rb").read()
    d = gzip.decompress(raw)
    assert d[:4] == b"BAM\x01"
    lt = struct.unpack("<i", d[4:8])[0]
    text = d[8:8 + lt]
    o = 8 + lt
    nref = struct.unpack("<i", d[o:o + 4])[0]
    o += 4
    refs = []
    for _ in range(nref):
        ln = struct.unpack("<i", d[o:o + 4])[0]
        refs.append((d[o + 4:o + 4 + ln], struct.unpack("<i", d[o + 4 + ln:o + 8 + ln])[0]))
        o += 8 + ln
    recs = []
    while o < len(d):
        bs = struct.unpack("<i", d[o:o + 4])[0]
        recs.append(d[o:o + 4 + bs])
        o += 4 + bs
    return raw, text, refs, recs


def check_bgzf(raw):
    """every member is a BGZF block (gzip header with the BC extra field, total size <= 64 KB); the file ends with the EOF marker"""
    o, n = 0, 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04" and raw[o + 12:o + 14] == b"BC"
        bsize = struct.unpack("<H", raw[o + 16:o + 18])[0] + 1
        assert bsize <= 0x10000
        o += bsize
        n += 1
    assert o == len(raw)
    assert raw[-28:] == bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    return n


def _header_lines(text):
    return [l for l in text.split(b"\n") if not l.startswith(b"@PG") and not l.startswith(b"@CO")]


CASES = [
    ("std", []),
    ("hard", []),
    ("se", []),
    ("std", ["--outSAMattributes", "NH", "HI", "AS", "nM", "NM", "MD", "jM", "jI", "MC", "XS", "--outSAMunmapped", "Within", "--outFilterMultimapNmax", "20"]),
    ("hard", ["--outSAMunmapped", "Within", "KeepPairs", "--outSAMattributes", "All", "--outSAMattrRGline", "ID:rg1", "SM:x", "--outSAMmode", "NoQS"]),
    ("hard", ["--alignEndsType", "EndToEnd", "--outSAMprimaryFlag", "AllBestScore", "--outSAMflagOR", "1024", "--outSAMattrIHstart", "0", "--outSAMmapqUnique", "60"]),
]


@pytest.mark.skipif(not os.path.exists(oc.REF_STAR), reason="oracle/_ref/STAR not built (needs /root/reference)")
@pytest.mark.parametrize("base,extra", CASES)
def test_bam_records_equal_the_reference(oracle, golden, tmp_path, base, extra):
    files = [os.path.join(golden, base + "_1.fq")] + ([os.path.join(golden, base + "_2.fq")] if base != "se" else [])
    outs = {}
    for tag, binary, thr in (("ref", oc.REF_STAR, 1), ("ora", oc.ORACLE_CLI, 3)):
        out = str(tmp_path / tag) + "/"
        os.makedirs(out)
        cmd = [binary, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn"] + files + ["--outFileNamePrefix", out, "--runThreadN", str(thr),
               "--outSAMtype", "BAM", "Unsorted"] + extra + (["--gpuChunkReads", "700"] if tag == "ora" else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, cwd=out)
        assert not os.path.exists(out + "Aligned.out.sam")
        outs[tag] = parse_bam(out + "Aligned.out.bam")
    assert check_bgzf(outs["ora"][0]) >= 2
    assert outs["ora"][2] == outs["ref"][2]
    assert _header_lines(outs["ora"][1]) == _header_lines(outs["ref"][1])
    assert len(outs["ora"][3]) == len(outs["ref"][3])
    for k, (x, y) in enumerate(zip(outs["ora"][3], outs["ref"][3])):
        assert x == y, "record %d differs" % k


def test_bam_record_count_and_names_match_the_sam_golden(oracle, golden, tmp_path):
    """Without the reference binary: the BAM of the std set has the same records (QNAME, FLAG, POS) as the committed SAM golden."""
    out = str(tmp_path) + "/"
    subprocess.check_call([oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq"),
                           "--outFileNamePrefix", out, "--runThreadN", "2", "--outSAMtype", "BAM", "Unsorted", "--outBAMcompression", "6"], stdout=subprocess.DEVNULL)
    raw, text, refs, recs = parse_bam(out + "Aligned.out.bam")
    check_bgzf(raw)
    sam = cf.sam_body(os.path.join(golden, "ref_std", "Aligned.out.sam"))
    assert len(recs) == len(sam)
    for rec, line in zip(recs, sam):
        f = line.split(b"\t")
        refid, pos, bmn, fn = struct.unpack("<iiII", rec[4:20])
        lname = bmn & 0xff
        assert rec[36:36 + lname - 1] == f[0] and (fn >> 16) == int(f[1]) and pos + 1 == int(f[3]) and ((bmn >> 8) & 0xff) == int(f[4])


def test_sharded_bam_merge(oracle, lib, golden, tmp_path):
    world = 2
    pre = str(tmp_path) + "/m_"
    args = ["--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "std_1.fq"), os.path.join(golden, "std_2.fq"), "--outSAMtype", "BAM", "Unsorted"]
    for r in range(world):
        subprocess.check_call([oc.ORACLE_CLI] + args + ["--outFileNamePrefix", pre + "shard%d." % r, "--gpuShardIndex", str(r), "--gpuShardCount", str(world)], stdout=subprocess.DEVNULL)
    argv = ["STAR"] + args + ["--outFileNamePrefix", pre]
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    lib.star_host_merge_shards.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_void_p]
    assert lib.star_host_merge_shards(len(argv), arr, world, None) == 0
    whole = str(tmp_path) + "/w_"
    subprocess.check_call([oc.ORACLE_CLI] + args + ["--outFileNamePrefix", whole], stdout=subprocess.DEVNULL)
    a, b = parse_bam(pre + "Aligned.out.bam"), parse_bam(whole + "Aligned.out.bam")
    check_bgzf(a[0])
    assert a[2] == b[2] and a[3] == b[3]


def test_sharded_sorted_bam_merge(oracle, lib, golden, tmp_path):
    """Coordinate-sorted BAM of a sharded run: every shard leaves its records + sort keys, the merge sorts the whole run; equal to the
    single-process file record for record (Within: the unmapped records come last, in read order, across shards)."""
    world = 3
    pre = str(tmp_path) + "/m_"
    args = ["--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "hard_1.fq"), os.path.join(golden, "hard_2.fq"),
            "--outSAMtype", "BAM", "Unsorted", "SortedByCoordinate", "--outSAMunmapped", "Within", "--runThreadN", "2"]
    for r in range(world):
        subprocess.check_call([oc.ORACLE_CLI] + args + ["--outFileNamePrefix", pre + "shard%d." % r, "--gpuShardIndex", str(r), "--gpuShardCount", str(world)], stdout=subprocess.DEVNULL)
    argv = ["STAR"] + args + ["--outFileNamePrefix", pre]
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    lib.star_host_merge_shards.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_void_p]
    assert lib.star_host_merge_shards(len(argv), arr, world, None) == 0
    whole = str(tmp_path) + "/w_"
    subprocess.check_call([oc.ORACLE_CLI] + args + ["--outFileNamePrefix", whole], stdout=subprocess.DEVNULL)
    for f in ("Aligned.out.bam", "Aligned.sortedByCoord.out.bam"):
        a, b = parse_bam(pre + f), parse_bam(whole + f)
        check_bgzf(a[0])
        assert a[2] == b[2] and len(a[3]) == len(b[3]) and a[3] == b[3], f


SORT_CASES = [
    ("std", []),
    ("hard", ["--outSAMunmapped", "Within"]),
    ("se", ["--outSAMunmapped", "Within", "--outSAMattributes", "NH", "HI", "AS", "nM", "NM", "MD"]),
    ("hard", ["--outSAMunmapped", "Within", "KeepPairs", "--outFilterMultimapNmax", "20", "--winAnchorMultimapNmax", "100"]),
]


@pytest.mark.skipif(not os.path.exists(oc.REF_STAR), reason="oracle/_ref/STAR not built (needs /root/reference)")
@pytest.mark.parametrize("base,extra", SORT_CASES)
def test_sorted_bam_equals_the_reference(oracle, golden, tmp_path, base, extra):
    """--outSAMtype BAM Unsorted SortedByCoordinate: both files; the sorted one must list the same records in the same order as the
    reference's bin-sorted file (coordinate, then read order; unmapped reads last in read order), header with SO:coordinate."""
    files = [os.path.join(golden, base + "_1.fq")] + ([os.path.join(golden, base + "_2.fq")] if base != "se" else [])
    outs = {}
    for tag, binary, thr in (("ref", oc.REF_STAR, 2), ("ora", oc.ORACLE_CLI, 3)):
        out = str(tmp_path / tag) + "/"
        os.makedirs(out)
        cmd = [binary, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn"] + files + ["--outFileNamePrefix", out, "--runThreadN", str(thr),
               "--outSAMtype", "BAM", "Unsorted", "SortedByCoordinate"] + extra + (["--gpuChunkReads", "500"] if tag == "ora" else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, cwd=out)
        outs[tag] = (parse_bam(out + "Aligned.sortedByCoord.out.bam"), parse_bam(out + "Aligned.out.bam"))
    srt_o, uns_o = outs["ora"]
    srt_r, uns_r = outs["ref"]
    check_bgzf(srt_o[0])
    assert srt_o[1].startswith(b"@HD\tVN:1.4\tSO:coordinate\n")
    assert _header_lines(srt_o[1]) == _header_lines(srt_r[1]) and srt_o[2] == srt_r[2]
    assert len(srt_o[3]) == len(srt_r[3])
    for k, (x, y) in enumerate(zip(srt_o[3], srt_r[3])):
        assert x == y, "sorted record %d differs" % k
    # the unsorted file next to it is unchanged by the extra output (the reference runs 2 threads here: compare as a multiset)
    assert sorted(uns_o[3]) == sorted(uns_r[3])
    keys = [struct.unpack("<II", r[4:12]) for r in srt_o[3]]
    assert keys == sorted(keys), "not coordinate-sorted"
