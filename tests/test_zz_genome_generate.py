"""--runMode genomeGenerate against index files written by the UNMODIFIED reference: the tiny genome with its GTF
(files in tests/golden/tiny.tar.gz, compared byte for byte) and a "torture" genome (tests/golden/genome.tar.gz, make_golden_genome.py:
repeats, reverse-complement copies, N runs, identical chromosomes, two FASTA files; compared by digest).

CPU: the host code (star_b200/csrc/host/genome_generate.cpp) with (a) the oracle's comparison sort and (b) the EMULATED kernels and round
loop of sa_build_impl.cuh (prefix doubling; cub's primitives replaced by std:: ones).  GPU: the drop-in CLI, i.e. star_gpu_sa_build.
"""
import hashlib
import json
import os
import subprocess
import tarfile

import pytest

import conftest as cf
import oracle_capi as oc

ROOT = cf.ROOT
EMUL = os.path.join(ROOT, "oracle", "_build", "libengine_emul.so")


@pytest.fixture(scope="module")
def torture(tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_genome")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "genome.tar.gz")) as t:
        t.extractall(d)
    return str(d / "genome")


def _generate(binary, cwd, out, args, env=None):
    os.makedirs(out, exist_ok=True)
    subprocess.check_call([binary, "--runMode", "genomeGenerate", "--genomeDir", out, "--outFileNamePrefix", out + "_log_"] + args, cwd=cwd, stdout=subprocess.DEVNULL, env=env)


def _check_torture(torture, out):
    for line in open(os.path.join(torture, "sha256.txt")):
        name, digest = line.split()
        data = open(os.path.join(out, name), "rb").read()
        if name == "genomeParameters.txt":
            data = data.split(b"\n", 1)[1]
        assert hashlib.sha256(data).hexdigest() == digest, name


def _check_tiny(golden, out):
    for name in ("Genome", "SA", "SAindex", "chrName.txt", "chrStart.txt", "chrLength.txt", "chrNameLength.txt", "sjdbInfo.txt", "sjdbList.out.tab", "sjdbList.fromGTF.out.tab",
                 "exonInfo.tab", "transcriptInfo.tab", "geneInfo.tab", "exonGeTrInfo.tab"):
        assert open(os.path.join(out, name), "rb").read() == open(os.path.join(golden, "idx", name), "rb").read(), name
    ours = open(os.path.join(out, "genomeParameters.txt")).read().split("\n", 1)[1]
    assert ours == open(os.path.join(golden, "idx", "genomeParameters.txt")).read().split("\n", 1)[1]


TINY_ARGS = ["--genomeFastaFiles", "genome.fa", "--sjdbGTFfile", "annot.gtf", "--sjdbOverhang", "99", "--genomeSAindexNbases", "7"]


@pytest.mark.parametrize("emulated", [False, True, "large"])
def test_generate_torture_genome_matches_reference(oracle, torture, tmp_path, emulated):
    """emulated = "large": the batched 64-bit path for texts beyond one sort (sa_build_large.cuh), forced with a small capacity so that
    round 0 runs in several bin runs and the doubling rounds in several batches of tied groups."""
    env = dict(os.environ, STAR_CLI_SJDB_EMUL=EMUL) if emulated else None
    if emulated == "large":
        env["STAR_B200_SA_LARGE_CAP"] = "20000"
    out = str(tmp_path / "idx") + "/"
    _generate(oc.ORACLE_CLI, torture, out, ["--genomeFastaFiles", "g1.fa", "g2.fa"] + json.load(open(os.path.join(torture, "args.json"))), env)
    _check_torture(torture, out)


@pytest.mark.parametrize("emulated", [False, True])
def test_generate_tiny_annotated_index_matches_reference(oracle, golden, tmp_path, emulated):
    env = dict(os.environ, STAR_CLI_SJDB_EMUL=EMUL) if emulated else None
    out = str(tmp_path / "idx") + "/"
    _generate(oc.ORACLE_CLI, golden, out, TINY_ARGS, env)
    _check_tiny(golden, out)


def test_generate_parameter_errors(oracle, golden, tmp_path):
    base = [oc.ORACLE_CLI, "--runMode", "genomeGenerate", "--genomeDir", str(tmp_path / "g"), "--outFileNamePrefix", str(tmp_path) + "/"]
    def rc(extra):
        return subprocess.run(base + extra, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=golden).returncode
    assert rc([]) == 102                                                                     # no --genomeFastaFiles
    assert rc(["--genomeFastaFiles", "missing.fa"]) == 104
    assert rc(["--genomeFastaFiles", "std_1.fq"]) == 104                                     # not FASTA
    assert rc(["--genomeFastaFiles", "genome.fa", "--sjdbOverhang", "50"]) == 104            # overhang without annotations
    assert rc(["--genomeFastaFiles", "genome.fa", "--sjdbGTFfile", "annot.gtf", "--sjdbOverhang", "0"]) == 104
    assert rc(["--genomeFastaFiles", "genome.fa", "--genomeSAsparseD", "2"]) == 102


@pytest.mark.gpu
def test_gpu_generate_matches_reference(lib, golden, torture, tmp_path):
    star = os.path.join(ROOT, "star_b200", "bin", "STAR")
    out = str(tmp_path / "idx_t") + "/"
    _generate(star, torture, out, ["--genomeFastaFiles", "g1.fa", "g2.fa"] + json.load(open(os.path.join(torture, "args.json"))))
    _check_torture(torture, out)
    out = str(tmp_path / "idx_tl") + "/"   # the batched 64-bit path, forced
    _generate(star, torture, out, ["--genomeFastaFiles", "g1.fa", "g2.fa"] + json.load(open(os.path.join(torture, "args.json"))), dict(os.environ, STAR_B200_SA_LARGE_CAP="20000"))
    _check_torture(torture, out)
    out = str(tmp_path / "idx_tiny") + "/"
    _generate(star, golden, out, TINY_ARGS)
    _check_tiny(golden, out)
