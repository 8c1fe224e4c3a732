#!/usr/bin/env python3
"""Regenerates tests/golden/tiny.tar.gz — run in the BUILD container only (needs oracle/_ref/STAR, i.e. /root/reference).

The reference repository holds no alignment tests or golden vectors, so the goldens are
outputs of the UNMODIFIED reference binary (oracle/_ref/STAR, built by oracle/Makefile.ref from
/root/reference/source) on seeded synthetic inputs (tools/synth.py):

  tiny/genome.fa, annot.gtf          3 chromosomes, 125 kb, 14 spliced transcripts          (seed 7 / 3)
  tiny/idx/                          reference --runMode genomeGenerate --sjdbOverhang 99 --genomeSAindexNbases 7
  tiny/std_{1,2}.fq                  2000 pairs 2x100, 0.5 % subst + indels + N + junk pairs (seed 1)
  tiny/hard_{1,2}.fq                 600 pairs 2x150, 5 % subst + indels + N                 (seed 5)
  tiny/se_1.fq                       500 single-end 100-mers                                 (seed 9)
  tiny/ref_<set>/Aligned.out.sam|SJ.out.tab|Log.final.out   reference outputs, --runThreadN 1, defaults
  tiny/ref_std_opts/...              reference outputs with non-default options (see OPTS below)

Nothing under /root/reference is read at test time; the tests unpack this archive.
"""
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402

STAR = os.path.join(ROOT, "oracle", "_ref", "STAR")
OPTS = ["--outSAMattributes", "NH", "HI", "AS", "nM", "NM", "MD", "jM", "jI", "MC", "--outSAMunmapped", "Within",
        "--outSAMstrandField", "intronMotif", "--outFilterMultimapNmax", "20", "--alignEndsType", "EndToEnd"]


def run(cmd, cwd):
    subprocess.check_call(cmd, cwd=cwd, stdout=subprocess.DEVNULL)


def main():
    tmp = tempfile.mkdtemp(prefix="golden_")
    d = os.path.join(tmp, "tiny")
    os.makedirs(d)
    chrs = synth.make_genome("tiny")
    trs = synth.make_annotation(chrs, "tiny")
    synth.write_fasta(chrs, os.path.join(d, "genome.fa"))
    synth.write_gtf(chrs, trs, os.path.join(d, "annot.gtf"))
    sets = {
        "std": dict(n_pairs=2000, read_len=100, mm=0.005, seed=1, indel=0.001, nrate=0.001, junk=0.02),
        "hard": dict(n_pairs=600, read_len=150, mm=0.05, seed=5, indel=0.002, nrate=0.002, junk=0.01),
        "se": dict(n_pairs=500, read_len=100, mm=0.01, seed=9, indel=0.001, nrate=0.001, junk=0.02),
    }
    for name, kw in sets.items():
        m1, m2 = synth.make_reads(chrs, trs, **kw)
        synth.write_fastq(m1, os.path.join(d, name + "_1.fq"))
        if name != "se":
            synth.write_fastq(m2, os.path.join(d, name + "_2.fq"))
    os.makedirs(os.path.join(d, "idx"))
    run([STAR, "--runMode", "genomeGenerate", "--genomeDir", "idx", "--genomeFastaFiles", "genome.fa", "--sjdbGTFfile", "annot.gtf",
         "--sjdbOverhang", "99", "--genomeSAindexNbases", "7", "--runThreadN", "4", "--outFileNamePrefix", "gen_"], d)
    for f in os.listdir(d):
        if f.startswith("gen_"):
            p = os.path.join(d, f)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)
    for f in os.listdir(os.path.join(d, "idx")):
        if f in ("Log.out",):
            os.remove(os.path.join(d, "idx", f))

    def align(tag, files, extra=()):
        out = os.path.join(d, "ref_" + tag)
        os.makedirs(out)
        run([STAR, "--genomeDir", "idx", "--readFilesIn"] + files + ["--outFileNamePrefix", "ref_" + tag + "/", "--runThreadN", "1"] + list(extra), d)
        for f in os.listdir(out):
            if f not in ("Aligned.out.sam", "SJ.out.tab", "Log.final.out"):
                p = os.path.join(out, f)
                shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)

    align("std", ["std_1.fq", "std_2.fq"])
    align("hard", ["hard_1.fq", "hard_2.fq"])
    align("se", ["se_1.fq"])
    align("std_opts", ["std_1.fq", "std_2.fq"], OPTS)
    dst = os.path.join(ROOT, "tests", "golden", "tiny.tar.gz")
    with tarfile.open(dst, "w:gz", compresslevel=9) as t:
        t.add(d, arcname="tiny")
    print("wrote", dst, os.path.getsize(dst), "bytes")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
