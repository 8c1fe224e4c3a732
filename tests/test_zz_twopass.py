"""On-the-fly junction insertion, --twopassMode Basic and --outFilterType BySJout (the S* scenarios) against outputs of the UNMODIFIED reference
(tests/golden/twopass.tar.gz, make_golden_twopass.py): Aligned.out.sam, SJ.out.tab, Log.final.out of both passes, sjdbInfo.txt /
sjdbList.out.tab and the bytes of the rebuilt Genome / SA / SAindex (--sjdbInsertSave All, compared by digest).

CPU: the repository's host code (star_b200/csrc/host/sjdb_insert.cpp + driver) with (a) the oracle's sequential restatement of the
two device steps and (b) the EMULATED CUDA kernels of sjdb_kernels.cuh (oracle/engine_emul.cpp through the test-only CLI).
GPU: the drop-in CLI, i.e. the real kernels behind star_gpu_sjdb_*.
"""
import json
import os
import subprocess

import pytest

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT
NAMES = ["A_novel", "B_annot", "C_files_hard", "D_annot_files_se", "E_insert_only", "F_gtf_insert", "G_gtf_files_twopass",
         "S1_bysjout", "S2_bysjout_annot_within", "S3_bysjout_se_filters", "S4_bysjout_twopass", "T_encode_twopass", "U_unmapped_fastx_bysjout", "Q_genecounts_bysjout", "R_transcriptome_sam", "R2_transcriptome_sam_bysjout_rg", "T2_encode_full",
         "V_clip_pe_all_outputs", "V2_clip_adapter_se", "V3_clip_mate_to_nothing"]


def _args(tp, golden, name):
    sc = json.load(open(os.path.join(tp, "scenarios.json")))[name]
    out = []
    for a in sc:
        if a.startswith("TP/"):
            out.append(os.path.join(tp, a[3:]))
        elif a in ("idx",) or a.endswith(".fq") or a.endswith(".gtf"):
            out.append(os.path.join(golden, a))
        else:
            out.append(a)
    return out


@pytest.mark.parametrize("name", NAMES)
def test_oracle_cli_twopass_matches_reference(oracle, golden, twopass_golden, tmp_path, name):
    out = str(tmp_path) + "/"
    subprocess.check_call([oc.ORACLE_CLI] + _args(twopass_golden, golden, name) + ["--outFileNamePrefix", out, "--runThreadN", "2"], stdout=subprocess.DEVNULL)
    cf.check_twopass_outputs(out, os.path.join(twopass_golden, name))


@pytest.mark.parametrize("name", ["B_annot", "C_files_hard"])
def test_emulated_sjdb_kernels_match_reference(oracle, golden, twopass_golden, tmp_path, name):
    """sjdb_search_kernel / sjdb_merge_sa_kernel compiled for the host: old junctions that move (B), two insertions in one run (C)."""
    out = str(tmp_path) + "/"
    env = dict(os.environ, STAR_CLI_SJDB_EMUL=os.path.join(ROOT, "oracle", "_build", "libengine_emul.so"))
    subprocess.check_call([oc.ORACLE_CLI] + _args(twopass_golden, golden, name) + ["--outFileNamePrefix", out, "--runThreadN", "2"], stdout=subprocess.DEVNULL, env=env)
    cf.check_twopass_outputs(out, os.path.join(twopass_golden, name))


def test_twopass_parameter_errors(oracle, golden, twopass_golden, tmp_path):
    """Parameters.cpp:779-825, 1019-1026 and sjdbPrepare.cpp:24-29: exit codes of the reference."""
    base = [oc.ORACLE_CLI, "--genomeDir", os.path.join(twopass_golden, "idx0"), "--readFilesIn", os.path.join(golden, "se_1.fq"), "--outFileNamePrefix", str(tmp_path) + "/"]
    def rc(extra):
        return subprocess.run(base + extra, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
    assert rc(["--twopass1readsN", "100"]) == 102                 # without --twopassMode
    assert rc(["--twopassMode", "Full"]) == 102
    assert rc(["--twopassMode", "Basic", "--twopass1readsN", "0"]) == 102
    assert rc(["--twopassMode", "Basic", "--sjdbOverhang", "0"]) == 102
    assert rc(["--twopassMode", "Basic", "--gpuShardCount", "2", "--gpuShardIndex", "0"]) == 102
    bad = os.path.join(str(tmp_path), "bad.tab")
    open(bad, "w").write("chrNope\t100\t200\t+\n")
    assert rc(["--sjdbFileChrStartEnd", bad]) == 104
    assert rc(["--sjdbFileChrStartEnd", os.path.join(str(tmp_path), "missing.tab")]) == 104


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_cli_twopass_matches_reference(lib, golden, twopass_golden, tmp_path, name):
    out = str(tmp_path) + "/"
    subprocess.check_call([os.path.join(ROOT, "star_b200", "bin", "STAR")] + _args(twopass_golden, golden, name) + ["--outFileNamePrefix", out, "--runThreadN", "2"],
                          stdout=subprocess.DEVNULL)
    cf.check_twopass_outputs(out, os.path.join(twopass_golden, name))


def test_housekeeping_parameters_are_accepted(oracle, golden, tmp_path):
    """Resource limits / temporary-file knobs of the reference have no effect here and must not break existing command lines;
    --readFilesPrefix, --outSAMorder PairedKeepInputOrder and --genomeLoad LoadAndKeep map onto what the GPU build does anyway."""
    out = str(tmp_path) + "/"
    cmd = [oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesPrefix", golden + "/", "--readFilesIn", "se_1.fq", "--outFileNamePrefix", out,
           "--limitBAMsortRAM", "1000000000", "--outBAMsortingThreadN", "2", "--limitOutSJcollapsed", "2000000", "--outSAMorder", "PairedKeepInputOrder",
           "--genomeLoad", "LoadAndKeep", "--runRNGseed", "5", "--outTmpKeep", "None", "--runThreadN", "2"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, cwd=str(tmp_path))
    ref = os.path.join(golden, "ref_se")
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(ref, "Aligned.out.sam"))
    assert "--limitBAMsortRAM is accepted and has no effect" in open(out + "Log.out").read()
    r = subprocess.run(cmd + ["--twopassMode", "Basic"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert r.returncode == 102 and "cannot be used with shared memory genome" in r.stderr


def test_outstd_streams_alignments_to_stdout(oracle, golden, tmp_path):
    """--outStd SAM | BAM_Unsorted | BAM_SortedByCoordinate (Parameters.cpp:385-391, 634-670): the alignments go to stdout, the messages the
    reference prints to stdout go to Log.std.out, the file is not created.  (Checked against the reference's own stdout when this was written.)"""
    import gzip
    base = [oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "se_1.fq"), "--runThreadN", "2"]
    ref = [l for l in open(os.path.join(golden, "ref_se", "Aligned.out.sam"), "rb").read().split(b"\n") if l and not l.startswith(b"@")]
    out = str(tmp_path / "a") + "/"
    r = subprocess.run(base + ["--outFileNamePrefix", out, "--outStd", "SAM"], stdout=subprocess.PIPE, check=True)
    assert [l for l in r.stdout.split(b"\n") if l and not l.startswith(b"@")] == ref
    assert r.stdout.startswith(b"@HD\tVN:1.4\n") and os.path.exists(out + "Log.std.out") and not os.path.exists(out + "Aligned.out.sam")
    assert b"started mapping" in open(out + "Log.std.out", "rb").read()
    out = str(tmp_path / "b") + "/"
    r = subprocess.run(base + ["--outFileNamePrefix", out, "--outStd", "BAM_Unsorted", "--outSAMtype", "BAM", "Unsorted", "SortedByCoordinate"], stdout=subprocess.PIPE, check=True)
    d = gzip.decompress(r.stdout)
    assert d[:4] == b"BAM\x01" and not os.path.exists(out + "Aligned.out.bam") and os.path.exists(out + "Aligned.sortedByCoord.out.bam")
    names = set(l.split(b"\t")[0] for l in ref)
    assert all(n in d for n in list(names)[:20])
    r = subprocess.run(base + ["--outFileNamePrefix", out, "--outStd", "Nonsense"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 102


def test_read_files_manifest(oracle, golden, tmp_path):
    """--readFilesManifest (Parameters_readFilesInit.cpp:96-137): files and read groups from a table; @RG header lines, RG tags only when asked for.
    (Equal to the reference's output when this was written; here: records equal the plain run's, header and tags as specified.)"""
    lines = open(os.path.join(golden, "se_1.fq")).read().split("\n")
    a, b = str(tmp_path / "sA.fq"), str(tmp_path / "sB.fq")
    open(a, "w").write("\n".join(lines[:200]) + "\n")
    open(b, "w").write("\n".join(lines[200:]))
    man = str(tmp_path / "man.tsv")
    open(man, "w").write("%s\t-\tID:grpA\tSM:x\n\n%s\t-\tgrpB\tSM:y\tPL:ill\n" % (a, b))
    out = str(tmp_path / "o") + "/"
    subprocess.check_call([oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesManifest", man, "--outFileNamePrefix", out, "--runThreadN", "2"], stdout=subprocess.DEVNULL)
    sam = open(out + "Aligned.out.sam").read()
    assert "@RG\tID:grpA\tSM:x\n@RG\tID:grpB\tSM:y\tPL:ill\n" in sam and "RG:Z:" not in sam
    assert cf.sam_body(out + "Aligned.out.sam") == cf.sam_body(os.path.join(golden, "ref_se", "Aligned.out.sam"))
    out = str(tmp_path / "p") + "/"
    subprocess.check_call([oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesManifest", man, "--outFileNamePrefix", out, "--outSAMattributes", "NH", "HI", "RG"], stdout=subprocess.DEVNULL)
    body = [l for l in open(out + "Aligned.out.sam").read().split("\n") if l and not l.startswith("@")]
    assert body[0].endswith("RG:Z:grpA") and body[-1].endswith("RG:Z:grpB")
    bad = str(tmp_path / "bad.tsv")
    open(bad, "w").write("x.fq\t-\n")
    assert subprocess.run([oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesManifest", bad, "--outFileNamePrefix", out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 104


@pytest.mark.gpu
def test_gpu_seed_search_lmax_matches_oracle_cli(lib, oracle, golden, tmp_path):
    """--seedSearchLmax on the GPU (the extra fixed-length search of seed_search_kernel) vs the same host code driven by the oracle
    (the oracle equals the reference for this option: tools/fuzz_mapping_params.py runs)."""
    outs = []
    for tag, binary in (("gpu", os.path.join(ROOT, "star_b200", "bin", "STAR")), ("ora", oc.ORACLE_CLI)):
        out = str(tmp_path / tag) + "/"
        os.makedirs(out)
        subprocess.check_call([binary, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "hard_1.fq"), os.path.join(golden, "hard_2.fq"),
                               "--outFileNamePrefix", out, "--runThreadN", "2", "--seedSearchLmax", "25", "--seedSearchStartLmax", "30"], stdout=subprocess.DEVNULL)
        outs.append(out)
    assert cf.sam_body(outs[0] + "Aligned.out.sam") == cf.sam_body(outs[1] + "Aligned.out.sam")
    assert open(outs[0] + "SJ.out.tab", "rb").read() == open(outs[1] + "SJ.out.tab", "rb").read()
    assert cf.log_counters(outs[0] + "Log.final.out") == cf.log_counters(outs[1] + "Log.final.out")


def test_quality_conversion_and_tlen_options(oracle, golden, tmp_path):
    """--outQSconversionAdd (readLoad.cpp:71-81) and --outSAMtlen (ReadAlign_alignBAM.cpp:84-88; both equal to the reference when written)."""
    out = str(tmp_path) + "/"
    base = [oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "se_1.fq"), "--outFileNamePrefix", out]
    subprocess.check_call(base + ["--outQSconversionAdd", "1"], stdout=subprocess.DEVNULL)
    a = cf.sam_body(out + "Aligned.out.sam")
    b = cf.sam_body(os.path.join(golden, "ref_se", "Aligned.out.sam"))
    assert len(a) == len(b) and all(x.split(b"\t")[:10] == y.split(b"\t")[:10] and x.split(b"\t")[10] == bytes(c + 1 for c in y.split(b"\t")[10]) for x, y in zip(a, b))
    assert subprocess.run(base + ["--outSAMtlen", "3"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 102


def test_parameters_file(oracle, golden, tmp_path):
    """--parametersFiles (Parameters.cpp:331-440): values from the file, overridden by the command line; equal to the reference when written."""
    pf = str(tmp_path / "pf.txt")
    open(pf, "w").write("# comment\noutSAMattributes NH HI AS nM XS\n\noutFilterMultimapNmax 5\nalignEndsType EndToEnd\n")
    out = str(tmp_path / "o") + "/"
    base = [oc.ORACLE_CLI, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "se_1.fq"), "--outFileNamePrefix", out]
    subprocess.check_call(base + ["--parametersFiles", pf], stdout=subprocess.DEVNULL)
    sam = open(out + "Aligned.out.sam").read()
    assert "XS:A:" in sam and not any("S\t" in l.split("\t")[5] + "\t" for l in sam.split("\n") if l and not l.startswith("@") and l.split("\t")[5] != "*")   # EndToEnd from the file
    subprocess.check_call(base + ["--parametersFiles", pf, "--alignEndsType", "Local"], stdout=subprocess.DEVNULL)   # the command line wins
    assert any("S" in l.split("\t")[5] for l in open(out + "Aligned.out.sam").read().split("\n") if l and not l.startswith("@"))
    open(pf, "a").write("alignEndsType Local\n")
    r = subprocess.run(base + ["--parametersFiles", pf], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 102 and "duplicate parameter" in r.stderr and "parametersFiles" in r.stderr


@pytest.mark.gpu
def test_gpu_engine_takes_empty_mates(lib, oracle, golden):
    """Empty mates of a pair (clipped to nothing before mapping) through the C-ABI on the GPU: equal to the oracle field by field."""
    import numpy as np
    import star_b200 as sb
    m1 = cf.read_fastq_seqs(os.path.join(golden, "std_1.fq"))[:64]
    m2 = cf.read_fastq_seqs(os.path.join(golden, "std_2.fq"))[:64]
    parts, off = [], [0]
    for i, (a, b) in enumerate(zip(m1, m2)):
        a, b = bytes(a), bytes(b)
        if i % 3 == 0:
            b = b""
        if i % 4 == 1:
            a = b""
        if i % 16 == 5:
            a, b = a[:19], b""
        if i % 16 == 7:
            a, b = b"", b""
        for s in (a, b):
            parts.append(s)
            off.append(off[-1] + len(s))
    seq = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    off = np.array(off, dtype=np.uint64)
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    oe = oc.OracleEngine(oracle, idx)
    res_o, al_o, _ = oe.map_chunk(seq, off, 64, 2)
    oe.close()
    eng = sb.Engine(lib, idx, max_reads=64)
    res_g, al_g, _ = eng.map_chunk(seq, off, 64, 2)
    eng.close()
    idx.close()
    diffs = oc.compare_outputs(res_o, al_o, res_g, al_g)
    assert not diffs, "\n".join(diffs[:10])
