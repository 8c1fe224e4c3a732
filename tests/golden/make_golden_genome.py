#!/usr/bin/env python3
"""Regenerates tests/golden/genome.tar.gz — run in the BUILD container only (needs oracle/_ref/STAR, i.e. /root/reference).

Inputs and reference digests for --runMode genomeGenerate:
  genome/g1.fa, g2.fa   seeded "torture" genome in two FASTA files: an exact 3 kb repeat, its reverse complement, N runs, a lowercase copy,
                        IUPAC codes, low-complexity runs, two IDENTICAL chromosomes (the suffix comparison runs to the chromosome end),
                        a chromosome whose length is a multiple of the bin size, a long reverse-complement copy
  genome/args.json      the genomeGenerate arguments (--genomeSAindexNbases 6 --genomeChrBinNbits 10)
  genome/sha256.txt     digests of the files the UNMODIFIED reference wrote: Genome, SA, SAindex, chrName.txt, chrStart.txt,
                        chrLength.txt, chrNameLength.txt, and genomeParameters.txt without its first (command line) line
The tiny genome of tiny.tar.gz is the second case: its reference-built index files are in that archive already.
"""
import hashlib
import json
import os
import random
import shutil
import subprocess
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAR = os.path.join(ROOT, "oracle", "_ref", "STAR")
ARGS = ["--genomeSAindexNbases", "6", "--genomeChrBinNbits", "10"]
FILES = ["Genome", "SA", "SAindex", "chrName.txt", "chrStart.txt", "chrLength.txt", "chrNameLength.txt", "genomeParameters.txt"]


def digest(path):
    data = open(path, "rb").read()
    if path.endswith("genomeParameters.txt"):
        data = data.split(b"\n", 1)[1]
    return hashlib.sha256(data).hexdigest()


def make_fasta(d):
    random.seed(11)
    rnd = lambda n: "".join(random.choice("ACGT") for _ in range(n))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rc = lambda s: "".join(comp.get(c, "N") for c in reversed(s))
    a = rnd(30000)
    seg = a[5000:8000]
    chr1 = a[:12000] + "N" * 137 + a[12000:] + seg + "NNNNN" + rc(seg) + rnd(500)
    chr2 = rnd(7000) + seg.lower() + "RYKM" + rnd(1000) + "A" * 300 + "AC" * 200 + rnd(777)
    chr3 = chr2
    chr4 = rnd(2048)
    chr5 = "N" * 50 + rnd(100) + "N" * 50
    chr6 = rc(chr1[:9000])

    def wr(fn, recs, w):
        with open(os.path.join(d, fn), "w") as f:
            for name, s in recs:
                f.write(">" + name + " some description\n")
                for i in range(0, len(s), w):
                    f.write(s[i:i + w] + "\n")
    wr("g1.fa", [("chrA", chr1), ("chrB", chr2), ("chrC", chr3)], 70)
    wr("g2.fa", [("chrD", chr4), ("chrE", chr5), ("chrF", chr6)], 61)


def main():
    tmp = tempfile.mkdtemp(prefix="golden_gen_")
    d = os.path.join(tmp, "genome")
    os.makedirs(d)
    make_fasta(d)
    ref = os.path.join(tmp, "ref")
    os.makedirs(ref)
    subprocess.check_call([STAR, "--runMode", "genomeGenerate", "--genomeDir", ref, "--genomeFastaFiles", "g1.fa", "g2.fa", "--runThreadN", "4",
                           "--outFileNamePrefix", os.path.join(tmp, "ref_")] + ARGS, cwd=d, stdout=subprocess.DEVNULL)
    with open(os.path.join(d, "sha256.txt"), "w") as f:
        for name in FILES:
            f.write("%s\t%s\n" % (name, digest(os.path.join(ref, name))))
    json.dump(ARGS, open(os.path.join(d, "args.json"), "w"))
    dst = os.path.join(ROOT, "tests", "golden", "genome.tar.gz")
    with tarfile.open(dst, "w:gz", compresslevel=9) as t:
        t.add(d, arcname="genome")
    print("wrote", dst, os.path.getsize(dst), "bytes")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
