#!/usr/bin/env python3
"""Regenerates tests/golden/twopass.tar.gz — run in the BUILD container only (needs oracle/_ref/STAR, i.e. /root/reference).

Golden outputs of the UNMODIFIED reference binary for on-the-fly junction insertion, --twopassMode Basic and
--outFilterType BySJout (the S* scenarios),
on the inputs of tiny.tar.gz:

  twopass/idx0/                     reference genomeGenerate WITHOUT annotation (--genomeSAindexNbases 7): every junction is novel
  twopass/sj_{half,dot,opp,shift}.tab   junction lists derived from tiny/idx/sjdbList.out.tab: every 2nd line; every 3rd with strand '.';
                                    every 4th on the opposite strand; every 5th moved by 3 bases (mostly non-canonical)
  twopass/scenarios.json            name -> argument list (relative to the unpacked tiny/ directory; TP = the unpacked twopass/ directory)
  twopass/<name>/                   Aligned.out.sam, SJ.out.tab, Log.final.out, _STARpass1/{SJ.out.tab,Log.final.out},
                                    _STARgenome/{sjdbInfo.txt,sjdbList.out.tab,sha256.txt}; sha256.txt = digests of the Genome, SA and
                                    SAindex files the reference wrote with --sjdbInsertSave All
"""
import hashlib
import json
import os
import shutil
import subprocess
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STAR = os.path.join(ROOT, "oracle", "_ref", "STAR")

SCENARIOS = {
    "A_novel": ["--genomeDir", "TP/idx0", "--readFilesIn", "std_1.fq", "std_2.fq", "--twopassMode", "Basic", "--sjdbInsertSave", "All"],
    "B_annot": ["--genomeDir", "idx", "--readFilesIn", "std_1.fq", "std_2.fq", "--twopassMode", "Basic", "--sjdbInsertSave", "All"],
    "C_files_hard": ["--genomeDir", "TP/idx0", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--twopassMode", "Basic", "--sjdbInsertSave", "All",
                     "--sjdbFileChrStartEnd", "TP/sj_half.tab", "TP/sj_dot.tab", "TP/sj_opp.tab", "TP/sj_shift.tab", "--sjdbOverhang", "80"],
    "D_annot_files_se": ["--genomeDir", "idx", "--readFilesIn", "se_1.fq", "--twopassMode", "Basic", "--sjdbInsertSave", "All",
                         "--sjdbFileChrStartEnd", "TP/sj_dot.tab", "TP/sj_opp.tab", "TP/sj_shift.tab", "--twopass1readsN", "300"],
    "E_insert_only": ["--genomeDir", "TP/idx0", "--readFilesIn", "std_1.fq", "std_2.fq", "--sjdbInsertSave", "All",
                      "--sjdbFileChrStartEnd", "TP/sj_half.tab", "TP/sj_opp.tab"],
    "S1_bysjout": ["--genomeDir", "TP/idx0", "--readFilesIn", "std_1.fq", "std_2.fq", "--outFilterType", "BySJout"],
    "S2_bysjout_annot_within": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--outFilterType", "BySJout", "--outSAMunmapped", "Within"],
    "S3_bysjout_se_filters": ["--genomeDir", "TP/idx0", "--readFilesIn", "se_1.fq", "--outFilterType", "BySJout", "--outSJfilterCountUniqueMin", "1", "1", "1", "1",
                              "--outSJfilterOverhangMin", "20", "10", "10", "10"],
    "S4_bysjout_twopass": ["--genomeDir", "TP/idx0", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--outFilterType", "BySJout", "--twopassMode", "Basic",
                           "--outSAMattributes", "NH", "HI", "AS", "nM", "XS", "--sjdbInsertSave", "All"],
    # the ENCODE long-RNA option set (STAR manual, "ENCODE options") with 2-pass on the hard reads
    "T_encode_twopass": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--outFilterType", "BySJout", "--outSAMattributes", "NH", "HI", "AS", "NM", "MD",
                         "--outFilterMultimapNmax", "20", "--outFilterMismatchNmax", "999", "--outFilterMismatchNoverReadLmax", "0.04", "--alignIntronMin", "20",
                         "--alignIntronMax", "1000000", "--alignMatesGapMax", "1000000", "--alignSJoverhangMin", "8", "--alignSJDBoverhangMin", "1", "--sjdbScore", "1",
                         "--outSAMunmapped", "Within", "--twopassMode", "Basic", "--sjdbInsertSave", "All"],
    "U_unmapped_fastx_bysjout": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--outReadsUnmapped", "Fastx", "--outFilterType", "BySJout",
                                 "--outSAMunmapped", "Within"],
    "Q_genecounts_bysjout": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--quantMode", "GeneCounts", "--outFilterType", "BySJout"],
    "R_transcriptome_sam": ["--genomeDir", "idx", "--readFilesIn", "std_1.fq", "std_2.fq", "--quantMode", "TranscriptomeSAM", "GeneCounts"],
    "R2_transcriptome_sam_bysjout_rg": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--quantMode", "TranscriptomeSAM", "--outFilterType", "BySJout",
                                        "--quantTranscriptomeSAMoutput", "BanSingleEnd", "--outSAMattrRGline", "ID:x", "SM:y", "--outSAMattributes", "NH", "HI", "AS", "nM", "RG", "MC"],
    # the complete ENCODE long-RNA command line (STAR manual / ENCODE pipeline): BySJout, both quant modes, header options, both BAM files
    "T2_encode_full": ["--genomeDir", "idx", "--readFilesIn", "std_1.fq", "std_2.fq", "--outFilterType", "BySJout", "--outSAMattributes", "NH", "HI", "AS", "NM", "MD",
                       "--outFilterMultimapNmax", "20", "--outFilterMismatchNmax", "999", "--outFilterMismatchNoverReadLmax", "0.04", "--alignIntronMin", "20",
                       "--alignIntronMax", "1000000", "--alignMatesGapMax", "1000000", "--alignSJoverhangMin", "8", "--alignSJDBoverhangMin", "1", "--sjdbScore", "1",
                       "--outSAMtype", "BAM", "Unsorted", "SortedByCoordinate", "--quantMode", "TranscriptomeSAM", "GeneCounts", "--outSAMunmapped", "Within",
                       "--outSAMstrandField", "intronMotif", "--outSAMheaderHD", "@HD", "VN:1.4", "SO:unsorted", "--outSAMheaderCommentFile", "TP/COfile.txt",
                       "--outSAMheaderPG", "@PG", "ID:x", "PN:y", "--limitBAMsortRAM", "10000000000"],
    # read clipping before mapping (--clip5pNbases / --clip3pNbases / 3' adapter): soft clips in the CIGARs, NM / MD / MC, both BAM files, quantification
    "V_clip_pe_all_outputs": ["--genomeDir", "idx", "--readFilesIn", "std_1.fq", "std_2.fq", "--clip5pNbases", "4", "9", "--clip3pNbases", "12", "6", "--outSAMtype", "BAM", "Unsorted",
                              "SortedByCoordinate", "--outSAMunmapped", "Within", "--quantMode", "TranscriptomeSAM", "GeneCounts", "--outReadsUnmapped", "Fastx",
                              "--outSAMattributes", "NH", "HI", "AS", "nM", "NM", "MD", "MC"],
    "V2_clip_adapter_se": ["--genomeDir", "idx", "--readFilesIn", "se_1.fq", "--clip3pAdapterSeq", "AGGTC", "--clip3pAdapterMMp", "0.2", "--clip3pAfterAdapterNbases", "2",
                           "--outSAMunmapped", "Within"],
    "V3_clip_mate_to_nothing": ["--genomeDir", "idx", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--clip3pNbases", "0", "200", "--clip3pAdapterSeq", "GATC", "polyA",
                                "--clip3pAdapterMMp", "0.1", "0.3", "--outSAMunmapped", "Within", "--outSAMattributes", "NH", "HI", "AS", "nM", "NM", "MD"],
    "F_gtf_insert": ["--genomeDir", "TP/idx0", "--readFilesIn", "std_1.fq", "std_2.fq", "--sjdbGTFfile", "annot.gtf", "--sjdbInsertSave", "All", "--sjdbOverhang", "99"],
    "G_gtf_files_twopass": ["--genomeDir", "TP/idx0", "--readFilesIn", "hard_1.fq", "hard_2.fq", "--sjdbGTFfile", "annot.gtf", "--sjdbFileChrStartEnd", "TP/sj_opp.tab",
                            "TP/sj_shift.tab", "--twopassMode", "Basic", "--sjdbInsertSave", "All"],
}
KEEP = ["Aligned.out.sam", "SJ.out.tab", "Log.final.out", "_STARpass1/SJ.out.tab", "_STARpass1/Log.final.out", "_STARgenome/sjdbInfo.txt",
        "_STARgenome/sjdbList.out.tab", "_STARgenome/sjdbList.fromGTF.out.tab", "_STARgenome/exonInfo.tab", "_STARgenome/transcriptInfo.tab", "_STARgenome/geneInfo.tab",
        "_STARgenome/exonGeTrInfo.tab", "Unmapped.out.mate1", "Unmapped.out.mate2", "ReadsPerGene.out.tab", "Aligned.toTranscriptome.out.bam", "Aligned.out.bam",
        "Aligned.sortedByCoord.out.bam"]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def main():
    tmp = tempfile.mkdtemp(prefix="golden_tp_")
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "tiny.tar.gz")) as t:
        t.extractall(tmp)
    tiny = os.path.join(tmp, "tiny")
    tp = os.path.join(tmp, "twopass")
    os.makedirs(os.path.join(tp, "idx0"))
    subprocess.check_call([STAR, "--runMode", "genomeGenerate", "--genomeDir", os.path.join(tp, "idx0"), "--genomeFastaFiles", "genome.fa",
                           "--genomeSAindexNbases", "7", "--runThreadN", "4", "--outFileNamePrefix", os.path.join(tmp, "gen0_")], cwd=tiny, stdout=subprocess.DEVNULL)
    os.remove(os.path.join(tp, "idx0", "Log.out"))
    rows = [l.split("\t") for l in open(os.path.join(tiny, "idx", "sjdbList.out.tab")).read().splitlines()]
    with open(os.path.join(tp, "sj_half.tab"), "w") as f:
        f.writelines("\t".join(r) + "\n" for i, r in enumerate(rows, 1) if i % 2 == 0)
    with open(os.path.join(tp, "sj_dot.tab"), "w") as f:
        f.writelines("\t".join(r[:3] + ["."]) + "\n" for i, r in enumerate(rows, 1) if i % 3 == 0)
    with open(os.path.join(tp, "sj_opp.tab"), "w") as f:
        f.writelines("\t".join(r[:3] + ["-" if r[3] == "+" else "+"]) + "\n" for i, r in enumerate(rows, 1) if i % 4 == 0)
    with open(os.path.join(tp, "sj_shift.tab"), "w") as f:
        f.writelines("\t".join([r[0], str(int(r[1]) + 3), str(int(r[2]) + 3), r[3]]) + "\n" for i, r in enumerate(rows, 1) if i % 5 == 0)
    with open(os.path.join(tp, "COfile.txt"), "w") as f:
        f.write("@CO\tLIBID:ENCLB175ZZZ\n\n@CO\tREFID:ENCFF001RGS\n")
    for name, args in SCENARIOS.items():
        out = os.path.join(tmp, "run_" + name) + "/"
        os.makedirs(out)
        a = [x.replace("TP/", tp + "/") for x in args]
        subprocess.check_call([STAR] + a + ["--outFileNamePrefix", out, "--runThreadN", "1"], cwd=tiny, stdout=subprocess.DEVNULL)
        dst = os.path.join(tp, name)
        for k in KEEP:
            if os.path.exists(out + k):
                os.makedirs(os.path.dirname(os.path.join(dst, k)), exist_ok=True)
                shutil.copy(out + k, os.path.join(dst, k))
        if os.path.exists(out + "_STARgenome/SA"):
            with open(os.path.join(dst, "_STARgenome", "sha256.txt"), "w") as f:
                for g in ("Genome", "SA", "SAindex"):
                    f.write("%s\t%s\n" % (g, sha(out + "_STARgenome/" + g)))
    with open(os.path.join(tp, "scenarios.json"), "w") as f:
        json.dump(SCENARIOS, f, indent=1)
    dst = os.path.join(ROOT, "tests", "golden", "twopass.tar.gz")
    with tarfile.open(dst, "w:gz", compresslevel=9) as t:
        t.add(tp, arcname="twopass")
    print("wrote", dst, os.path.getsize(dst), "bytes")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
