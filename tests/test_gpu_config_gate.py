"""The bit-identity gate of BASELINE.json configs[0] at its real size, on the GPU (last bullet).

Genome: the chr21-sized synthetic stand-in (46.7 Mb, 3 chromosomes, repeat families, N block, 1352 spliced transcripts -> sjdb; there is no
real chr21 on the box), index built on the box by the UNMODIFIED reference's genomeGenerate.  The drop-in command line star_b200/bin/STAR
(CUDA engine) and oracle/_ref/STAR --runThreadN 1 map the same FASTQ files; the SAM records, SJ.out.tab and the integer counters of
Log.final.out must be byte-equal:
  * 100 k pairs 2x100 at 0.5 % substitutions (configs[0]),
  * 20 k pairs 2x150 at 5 % substitutions (configs[2] shape: long recursion trees, pool / task caps, overflow tiers),
and the engine's records for the same reads equal the oracle's field by field through the C-ABI.  The heavy tail that drives the
design (windows with 35-45 seeds from the repeat families, tens of thousands of recursion nodes per read, bump pools with per-chunk
offsets) only exists at this size.  Also here: this repository's GPU genomeGenerate reproduces the reference's index files of this genome.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import conftest as cf

pytestmark = pytest.mark.gpu

ROOT = cf.ROOT
REF = os.path.join(ROOT, "oracle", "_ref", "STAR")
OURS = os.path.join(ROOT, "star_b200", "bin", "STAR")


@pytest.fixture(scope="module")
def chr21(tmp_path_factory):
    """Work directory with genome.fa, annot.gtf and idx/ (reference genomeGenerate); shared with bench.py --preset chr21."""
    import bench
    import synth
    wd = os.path.join(os.environ.get("STAR_B200_BENCH_DIR", "/tmp/star_b200_bench"), "chr21")
    os.makedirs(wd, exist_ok=True)
    chrs, trs, idx, _ = bench.prepare_genome(wd, "chr21")
    return {"dir": wd, "chrs": chrs, "trs": trs, "idx": idx, "synth": synth}


def _reads(c, n, read_len, mm, seed, tag):
    s = c["synth"]
    m1, m2 = s.make_reads(c["chrs"], c["trs"], n, read_len=read_len, mm=mm, seed=seed)
    f1, f2 = os.path.join(c["dir"], tag + "_1.fq"), os.path.join(c["dir"], tag + "_2.fq")
    s.write_fastq(m1, f1)
    s.write_fastq(m2, f2)
    return m1, m2, f1, f2


def _run(binary, idx, f1, f2, out, extra=()):
    os.makedirs(out, exist_ok=True)
    subprocess.check_call([binary, "--genomeDir", idx, "--readFilesIn", f1, f2, "--outFileNamePrefix", out + "/"] + list(extra), stdout=subprocess.DEVNULL, timeout=1500)


@pytest.mark.parametrize("name,n,read_len,mm", [("std100", 100_000, 100, 0.005), ("hard150", 20_000, 150, 0.05)])
def test_cli_equals_reference_at_config_size(lib, chr21, tmp_path, name, n, read_len, mm):
    _, _, f1, f2 = _reads(chr21, n, read_len, mm, 77, "gate_" + name)
    ours, ref = str(tmp_path / "ours"), str(tmp_path / "ref")
    _run(OURS, chr21["idx"], f1, f2, ours, ["--runThreadN", "16"])
    _run(REF, chr21["idx"], f1, f2, ref, ["--runThreadN", "1"])
    a, b = cf.sam_body(ours + "/Aligned.out.sam"), cf.sam_body(ref + "/Aligned.out.sam")
    assert len(a) == len(b)
    bad = [i for i in range(len(a)) if a[i] != b[i]]
    assert not bad, "%d of %d SAM records differ, first:\n%s\n%s" % (len(bad), len(a), a[bad[0]], b[bad[0]])
    assert open(ours + "/SJ.out.tab", "rb").read() == open(ref + "/SJ.out.tab", "rb").read()
    assert cf.log_counters(ours + "/Log.final.out") == cf.log_counters(ref + "/Log.final.out")


@pytest.mark.parametrize("name,n,read_len,mm", [("std100", 100_000, 100, 0.005), ("hard150", 20_000, 150, 0.05)])
def test_engine_equals_oracle_at_config_size(lib, oracle, chr21, name, n, read_len, mm):
    """Through the C-ABI, one chunk: every field of every record, plus the work counters the roofline numerator is built from."""
    import oracle_capi as oc
    import star_b200 as sb
    m1, m2, _, _ = _reads(chr21, n, read_len, mm, 78, "gate2_" + name)
    seq, off, n_, nm = sb.pack_reads([m1, m2])
    index = sb.Index(lib, chr21["idx"])
    try:
        eng = sb.Engine(lib, index, max_reads=n_)
        res_g, al_g, st_g = eng.map_chunk(seq, off, n_, nm)
        eng.close()
        oe = oc.OracleEngine(oracle, index)
        res_o, al_o, st_o = oe.map_chunk(seq, off, n_, nm)
        oe.close()
    finally:
        index.close()
    diffs = oc.compare_outputs(res_o, al_o, res_g, al_g)
    assert not diffs, "\n".join(diffs[:20])
    for k in ("mmp_searches", "mmp_sai_words", "sa_enumerated"):
        assert getattr(st_g, k) == getattr(st_o, k), k


def test_gpu_generate_equals_reference_index_at_config_size(lib, chr21, tmp_path):
    """This repository's --runMode genomeGenerate (GPU suffix sort, both the 32-bit path and the batched 64-bit path that GRCh38 takes)
    writes the reference's Genome / SA / SAindex / junction files for the chr21-sized genome byte for byte."""
    ref = chr21["idx"]

    def digest(p):
        h = hashlib.sha256()
        with open(p, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        return h.hexdigest()
    for tag, env in (("small_path", {}), ("large_path", {"STAR_B200_SA_LARGE_CAP": "30000000"})):
        out = str(tmp_path / tag)
        os.makedirs(out)
        subprocess.check_call([OURS, "--runMode", "genomeGenerate", "--genomeDir", out, "--genomeFastaFiles", os.path.join(chr21["dir"], "genome.fa"),
                               "--sjdbGTFfile", os.path.join(chr21["dir"], "annot.gtf"), "--sjdbOverhang", "99", "--genomeSAindexNbases", "11", "--runThreadN", "16",
                               "--outFileNamePrefix", out + "_log_"], stdout=subprocess.DEVNULL, env=dict(os.environ, **env), timeout=1500)
        for f in ("Genome", "SA", "SAindex", "chrStart.txt", "chrLength.txt", "chrName.txt", "sjdbInfo.txt", "sjdbList.out.tab", "exonInfo.tab", "transcriptInfo.tab"):
            assert digest(os.path.join(out, f)) == digest(os.path.join(ref, f)), (tag, f)
