"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/star_b200.h declares, the host logic works,
and the engine refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import conftest as cf

ROOT = cf.ROOT


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "star_b200.h")).read()
    names = set(re.findall(r"\b(star_[a-z_0-9]+)\s*\(", hdr))
    names = {n for n in names if not n.endswith("_t")}
    assert {"star_gpu_init", "star_gpu_map_chunk", "star_gpu_destroy", "star_gpu_last_error", "star_cli_main", "star_index_load"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), "missing export: " + n


def test_abi_struct_sizes(lib):
    from star_b200 import capi
    lib.star_abi_sizeof.restype = C.c_size_t
    assert lib.star_abi_sizeof(0) == C.sizeof(capi.Params)
    assert lib.star_abi_sizeof(1) == C.sizeof(capi.IndexView)
    assert lib.star_abi_sizeof(2) == C.sizeof(capi.ReadBatch)
    assert lib.star_abi_sizeof(3) == capi.ALIGN_DTYPE.itemsize
    assert lib.star_abi_sizeof(4) == capi.RESULT_DTYPE.itemsize
    assert lib.star_abi_sizeof(5) == C.sizeof(capi.AlignBatch)
    assert lib.star_abi_sizeof(6) == C.sizeof(capi.ChunkStats)


def test_default_params_match_reference_defaults(lib):
    import star_b200 as sb
    p = sb.default_params(lib)
    # reference source/parametersDefault
    assert (p.seedSearchStartLmax, p.seedMapMin, p.seedSplitMin, p.seedMultimapNmax, p.seedPerReadNmax, p.seedPerWindowNmax) == (50, 5, 12, 10000, 1000, 50)
    assert (p.winAnchorMultimapNmax, p.winBinNbits, p.winAnchorDistNbins, p.winFlankNbins) == (50, 16, 9, 4)
    assert (p.alignIntronMin, p.alignSJoverhangMin, p.alignSJDBoverhangMin) == (21, 5, 3)
    assert list(p.alignSJstitchMismatchNmax) == [0, -1, 0, 0]
    assert (p.scoreGap, p.scoreGapNoncan, p.scoreGapGCAG, p.scoreGapATAC, p.scoreDelOpen, p.scoreInsBase, p.sjdbScore) == (0, -8, -4, -8, -2, -2, 2)
    assert p.scoreGenomicLengthLog2scale == -0.25 and p.outFilterMismatchNoverLmax == 0.3
    assert (p.outFilterMultimapNmax, p.outFilterMismatchNmax, p.outFilterMultimapScoreRange) == (10, 10, 1)
    assert p.outSAMmultNmax == 2 ** 64 - 1


def test_index_loader_reads_star_genome_dir(lib, golden):
    import star_b200 as sb
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    v = idx.view.contents
    assert v.GstrandBit == 32 and v.gSAindexNbases == 7 and v.nChrReal == 3
    assert v.nGenome == os.path.getsize(os.path.join(golden, "idx", "Genome"))
    assert v.nSA == os.path.getsize(os.path.join(golden, "idx", "SA")) * 8 // 33
    assert v.sjdbN > 0 and v.sjdbOverhang == 99 and v.sjdbLength == 199
    # window geometry derived at load time (Genome_genomeLoad.cpp:382-410)
    assert idx.params.winBinChrNbits == 2 and idx.params.winBinN == v.nGenome // 65536 + 1
    idx.close()
    with pytest.raises(sb.StarError):
        sb.Index(lib, "/nonexistent/genomeDir")


def test_engine_fails_loudly_without_gpu(lib, golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import star_b200 as sb
    idx = sb.Index(lib, os.path.join(golden, "idx"))
    with pytest.raises(sb.StarError) as e:
        sb.Engine(lib, idx, 16)
    assert "no CPU fallback" in str(e.value)
    idx.close()


def test_cli_rejects_bad_parameters(lib, golden, tmp_path):
    star = os.path.join(ROOT, "star_b200", "bin", "STAR")
    base = [star, "--genomeDir", os.path.join(golden, "idx"), "--readFilesIn", os.path.join(golden, "se_1.fq"), "--outFileNamePrefix", str(tmp_path) + "/"]
    r = subprocess.run(base + ["--noSuchParameter", "1"], capture_output=True, text=True)
    assert r.returncode == 102 and "unrecognized parameter name" in r.stderr
    r = subprocess.run(base + ["--soloType", "CB_UMI_Simple"], capture_output=True, text=True)
    assert r.returncode == 102 and "outside the scope" in r.stderr
    r = subprocess.run(base + ["--outSAMtype", "BAM", "SortedByName"], capture_output=True, text=True)
    assert r.returncode == 102 and "unknown value for the word 2 of outSAMtype" in r.stderr
    r = subprocess.run(base + ["--outFilterType", "BySJout", "--gpuShardIndex", "0", "--gpuShardCount", "2"], capture_output=True, text=True)
    assert r.returncode == 102 and "needs the junctions of all shards" in r.stderr   # (star_b200.dist runs the phases)
    r = subprocess.run(base + ["--outSAMtype", "BAM"], capture_output=True, text=True)
    assert r.returncode == 102 and "missing BAM option" in r.stderr
    r = subprocess.run([star, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "2.7.11b"


def test_product_does_not_reference_the_oracle():
    """The product path must not include, link or import anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "star_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in txt.split("\n"):
                    if "#include" in line or line.strip().startswith(("import ", "from ")):
                        assert "oracle" not in line, (f, line)
                assert "star_oracle" not in txt, f
    so = os.path.join(ROOT, "star_b200", "lib", "libstar_b200.so")
    if os.path.exists(so):
        out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
        assert "star_oracle" not in out


def test_pack_reads_layout():
    import star_b200 as sb
    m1 = np.frombuffer(b"ACGTACGTAAAA", dtype=np.uint8).reshape(3, 4)
    m2 = np.frombuffer(b"TTTTGGGGCCCC", dtype=np.uint8).reshape(3, 4)
    seq, off, n, nm = sb.pack_reads([m1, m2])
    assert n == 3 and nm == 2 and list(off) == [0, 4, 8, 12, 16, 20, 24]
    assert bytes(seq[:8]) == b"ACGTTTTT"
    seq2, off2, n2, nm2 = sb.pack_reads([[b"ACG", b"AC"], [b"T", b"GGGG"]])
    assert list(off2) == [0, 3, 4, 6, 10] and bytes(seq2) == b"ACGTACGGGG"
