// genome_generate.cpp — `STAR --runMode genomeGenerate`: FASTA -> Genome, SA, SAindex (+ junction inserts).
//
// Host side of reference source/Genome_genomeGenerate.cpp:98-415: FASTA scan and chromosome padding (genomeScanFastaFiles.cpp:5-103),
// chr*.txt (writeChrInfo, :417-432), SAindex (genomeSAindex.cpp:6-220), genomeParameters.txt (genomeParametersWrite.cpp:4-46) and the
// junction insertion shared with the mapping stage (sjdb_insert.cpp).  The suffix sort — hours of qsort over 16-mer buckets in the
// reference (:213-330) — is the device step behind star_gpu_sa_build (star_b200/csrc/engine/sa_build.cu).
#include <sys/stat.h>

#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "host.h"

namespace starhost {

namespace {

struct PackedRW {   // PackedArray.h:24-32, PackedArray.cpp:17-25
    uint8_t* a;
    uint32_t bits;
    uint64_t mask;
    PackedRW(uint8_t* p, uint32_t b) : a(p), bits(b), mask(~0ULL >> (64 - b)) {}
    uint64_t get(uint64_t i) const { const uint64_t b = i * bits; uint64_t w; memcpy(&w, a + b / 8, 8); return (w >> (b % 8)) & mask; }
    void set(uint64_t i, uint64_t x) {
        const uint64_t b = i * bits, S = b % 8;
        uint64_t w;
        memcpy(&w, a + b / 8, 8);
        w = (w & ~(mask << S)) | (x << S);
        memcpy(a + b / 8, &w, 8);
    }
};

// funCalcSAiFromSA, SuffixArrayFuns.cpp:353-395: prefix code of the first L bases of SA row iSA; iL4 = offset of the first base > 3 (-1: none)
uint64_t calcSAiFromSA(const uint8_t* G, const PackedRW& SA, uint64_t nGenome, uint32_t GstrandBit, uint64_t iSA, int L, int& iL4) {
    uint64_t SAstr = SA.get(iSA);
    const bool dirG = (SAstr >> GstrandBit) == 0;
    SAstr &= ~(1ULL << GstrandBit);
    iL4 = -1;
    uint64_t saind = 0;
    for (int ii = 0; ii < L; ii++) {
        const uint8_t g = dirG ? G[SAstr + ii] : G[nGenome - 1 - SAstr - ii];
        if (g > 3) { iL4 = ii; return saind << (2 * (L - ii)); }
        saind = (saind << 2) + (dirG ? g : 3 - g);
    }
    return saind;
}

}  // namespace

int genomeGenerate(HostParams& P, const star_engine_vtbl_t* eng, std::ostream& logMain, std::string& err) {
    const std::string& gDir = P.genomeDir;
    // Genome_genomeGenerate.cpp:113-133
    const bool annot = P.sjdbFileChrStartEnd[0] != "-" || P.sjdbGTFfile != "-";
    uint64_t sjdbOverhang = P.sjdbOverhang;
    if (annot && (long long)sjdbOverhang <= 0) {
        err = "EXITING because of FATAL INPUT PARAMETER ERROR: for generating genome with annotations (--sjdbFileChrStartEnd or --sjdbGTFfile options)\nyou need to specify >0 --sjdbOverhang\nSOLUTION: re-run genome generation specifying non-zero --sjdbOverhang, which ideally should be equal to OneMateLength-1, or could be chosen generically as ~100\n";
        return STAR_EXIT_INPUT_FILES;
    }
    if (!annot) {
        if (P.userSet.count("sjdbOverhang") && sjdbOverhang > 0) {
            err = "EXITING because of FATAL INPUT PARAMETER ERROR: when generating genome without annotations (--sjdbFileChrStartEnd or --sjdbGTFfile options)\ndo not specify >0 --sjdbOverhang\nSOLUTION: re-run genome generation without --sjdbOverhang option\n";
            return STAR_EXIT_INPUT_FILES;
        }
        sjdbOverhang = 0;
    }
    const uint64_t sjdbLength = sjdbOverhang == 0 ? 0 : 2 * sjdbOverhang + 1;
    { time_t t; time(&t); char b[100]; strftime(b, 80, "%b %d %H:%M:%S", localtime(&t)); std::cout << b << " ... starting to generate Genome files\n" << std::flush; }

    // ---- genomeScanFastaFiles.cpp:5-103: chromosomes start at multiples of 2^genomeChrBinNbits, gaps are code 5
    LoadedIndex idx;
    const uint64_t binN = 1ULL << P.genomeChrBinNbits;
    const size_t PAD = 256;
    std::vector<uint8_t>& Gs = idx.Gstore;
    Gs.assign(PAD, 5);
    uint64_t N = 0;
    auto padTo = [&](uint64_t n) { if (Gs.size() < PAD + n) Gs.resize(PAD + n, 5); };
    for (const std::string& fn : P.genomeFastaFiles) {
        std::ifstream in(fn);
        if (!in.good()) { err = "EXITING because of INPUT ERROR: could not open genomeFastaFile: " + fn + "\n"; return STAR_EXIT_INPUT_FILES; }
        const int cc = in.peek();
        if (!in.good()) { err = "EXITING because of INPUT ERROR: could not read from genomeFastaFile: " + fn + "\n"; return STAR_EXIT_INPUT_FILES; }
        if (cc != '>') {
            err = "EXITING because of INPUT ERROR: the file format of the genomeFastaFile: " + fn + " is not fasta: the first character is '" + std::string(1, (char)cc) + "' (" + std::to_string(cc) +
                  "), not '>'.\n Solution: check formatting of the fasta file. Make sure the file is uncompressed (unzipped).\n";
            return STAR_EXIT_INPUT_FILES;
        }
        std::string line;
        while (!in.eof()) {
            std::getline(in, line);
            if (!line.empty() && line[0] == '>') {
                std::istringstream ls(line);
                ls.ignore(1, ' ');
                std::string name;
                ls >> name;
                idx.chrName.push_back(name);
                if (!idx.chrStart.empty()) idx.chrLength.push_back(N - idx.chrStart.back());
                if (N > 0) N = ((N + 1) / binN + 1) * binN;
                idx.chrStart.push_back(N);
                padTo(N);
                logMain << fn << " : chr # " << idx.chrStart.size() - 1 << "  \"" << name << "\" chrStart: " << N << "\n";
            } else {   // convertNucleotidesToNumbersRemoveControls (SequenceFuns.cpp:170-192): control characters are dropped from the COUNT only
                padTo(N + line.size());
                uint64_t kept = 0;
                for (size_t jj = 0; jj < line.size(); jj++) {
                    const int c = (unsigned char)line[jj];
                    uint8_t v;
                    switch (c) {
                        case 'A': case 'a': v = 0; break;
                        case 'C': case 'c': v = 1; break;
                        case 'G': case 'g': v = 2; break;
                        case 'T': case 't': v = 3; break;
                        default: if (c < 32) continue; v = 4;
                    }
                    Gs[PAD + N + jj] = v;
                    kept++;
                }
                N += kept;
            }
        }
    }
    if (idx.chrStart.empty()) { err = "EXITING because of INPUT ERROR: no sequences in --genomeFastaFiles\n"; return STAR_EXIT_INPUT_FILES; }
    idx.chrLength.push_back(N - idx.chrStart.back());
    N = ((N + 1) / binN + 1) * binN;
    const uint32_t nChr = (uint32_t)idx.chrName.size();
    idx.chrStart.push_back(N);
    const uint64_t nGenome = N;
    Gs.resize(PAD + nGenome, 5);
    Gs.resize(PAD + nGenome + PAD, 5);
    uint64_t nGenomeTrue = 0;
    logMain << "Chromosome sequence lengths: \n";
    for (uint32_t i = 0; i < nChr; i++) { nGenomeTrue += idx.chrLength[i]; logMain << idx.chrName[i] << "\t" << idx.chrLength[i] << "\n"; }
    logMain << "Genome sequence total length = " << nGenomeTrue << "\nGenome size with padding = " << nGenome << "\n";
    if ((double)P.genomeSAindexNbases > std::log2((double)nGenomeTrue) / 2 - 1)
        logMain << "!!!!! WARNING: --genomeSAindexNbases " << P.genomeSAindexNbases << " is too large for the genome size=" << nGenomeTrue
                << ", which may cause seg-fault at the mapping step. Re-run genome generation with recommended --genomeSAindexNbases " << int(std::log2((double)nGenomeTrue) / 2 - 1) << "\n";
    mkdir(gDir.c_str(), 0755);
    {   // writeChrInfo, Genome_genomeGenerate.cpp:417-432
        std::ofstream cn(gDir + "chrName.txt"), cs(gDir + "chrStart.txt"), cl(gDir + "chrLength.txt"), cnl(gDir + "chrNameLength.txt");
        if (!cn.good()) { err = "EXITING because of fatal ERROR: could not create output file " + gDir + "chrName.txt\nSOLUTION: check the path and permissions of --genomeDir\n"; return STAR_EXIT_PARAMETER; }
        for (uint32_t i = 0; i < nChr; i++) {
            cn << idx.chrName[i] << "\n"; cs << idx.chrStart[i] << "\n"; cl << idx.chrLength[i] << "\n"; cnl << idx.chrName[i] << "\t" << idx.chrLength[i] << "\n";
        }
        cs << idx.chrStart[nChr] << "\n";
    }
    // ---- suffix array: every position of G + reverse complement that starts with a base (Genome_genomeGenerate.cpp:178-330)
    uint64_t nSA = 0;
    for (uint64_t i = 0; i < nGenome; i++) nSA += Gs[PAD + i] < 4;
    nSA *= 2;
    uint32_t GstrandBit = (uint32_t)std::floor(std::log((double)(nGenome + P.limitSjdbInsertNsj * sjdbLength)) / std::log(2.0)) + 1;
    if (GstrandBit < 32) GstrandBit = 32;
    logMain << "Estimated genome size with padding and SJs: total=genome+SJ=" << nGenome + P.limitSjdbInsertNsj * sjdbLength << " = " << nGenome << " + " << P.limitSjdbInsertNsj * sjdbLength
            << "\nGstrandBit=" << GstrandBit << "\nNumber of SA indices: " << nSA << "\n";
    if (nSA == 0) { err = "EXITING because of INPUT ERROR: the genome contains no A/C/G/T bases\n"; return STAR_EXIT_INPUT_FILES; }
    const uint64_t nSAbyte = (nSA - 1) * (GstrandBit + 1) / 8 + 8;
    idx.SAstore.assign(nSAbyte + 16, 0);
    { time_t t; time(&t); char b[100]; strftime(b, 80, "%b %d %H:%M:%S", localtime(&t)); std::cout << b << " ... starting to sort Suffix Array. This may take a long time...\n" << std::flush; }
    int rc = eng->sa_build(P.gpuDevice, Gs.data() + PAD, nGenome, GstrandBit, nSA, idx.SAstore.data(), nSAbyte);
    if (rc) { err = std::string("EXITING because of FATAL ERROR: suffix array generation failed: ") + eng->last_error() + "\n"; return rc; }

    // ---- SAindex, genomeSAindex.cpp:6-220 (the jump + bisect walk over runs of equal prefixes)
    { time_t t; time(&t); char b[100]; strftime(b, 80, "%b %d %H:%M:%S", localtime(&t)); std::cout << b << " ... generating Suffix Array index\n" << std::flush; }
    const uint32_t Lmax = (uint32_t)P.genomeSAindexNbases;
    idx.genomeSAindexStart.assign(Lmax + 1, 0);
    for (uint32_t i = 1; i <= Lmax; i++) idx.genomeSAindexStart[i] = idx.genomeSAindexStart[i - 1] + (1ULL << (2 * i));
    const uint64_t nSAi = idx.genomeSAindexStart[Lmax];
    const uint64_t nSAibyte = (nSAi - 1) * (GstrandBit + 3) / 8 + 8;
    idx.SAistore.assign(nSAibyte + 16, 0);
    {
        const uint8_t* G = Gs.data() + PAD;
        PackedRW SA(idx.SAstore.data(), GstrandBit + 1), SAi(idx.SAistore.data(), GstrandBit + 3);
        const uint64_t nC = 1ULL << (GstrandBit + 1), absentC = 1ULL << (GstrandBit + 2);
        const uint64_t* start = idx.genomeSAindexStart.data();
        std::vector<uint64_t> ind0(Lmax, ~0ULL);   // last prefix seen per length (-1: none yet)
        const uint64_t isaStep = nSA / (1ULL << (2 * Lmax)) + 1;
        uint64_t isa = 0;
        int iL4;
        uint64_t indFull = calcSAiFromSA(G, SA, nGenome, GstrandBit, isa, Lmax, iL4);
        while (isa <= nSA - 1) {
            for (uint32_t iL = 0; iL < Lmax; iL++) {
                const uint64_t indPref = indFull >> (2 * (Lmax - 1 - iL));
                if ((int)iL == iL4) {   // a base > 3 inside the prefix: flag the last present prefix of every longer length
                    for (uint32_t iL1 = iL; iL1 < Lmax; iL1++) SAi.set(start[iL1] + ind0[iL1], SAi.get(start[iL1] + ind0[iL1]) | nC);
                    break;
                }
                if (indPref > ind0[iL] || isa == 0) {
                    SAi.set(start[iL] + indPref, isa);
                    for (uint64_t ii = ind0[iL] + 1; ii < indPref; ii++) SAi.set(start[iL] + ii, isa | absentC);
                    ind0[iL] = indPref;
                } else if (indPref < ind0[iL]) { err = "BUG: next index is smaller than previous, EXITING\n"; return STAR_EXIT_INPUT_FILES; }
            }
            // funSAiFindNextIndex: first row whose (prefix, N offset) differs — steps of isaStep, then bisection
            const uint64_t indPrev = indFull;
            const int iL4prev = iL4;
            isa += isaStep;
            while (isa < nSA && (indFull = calcSAiFromSA(G, SA, nGenome, GstrandBit, isa, Lmax, iL4)) == indPrev && iL4 == iL4prev) isa += isaStep;
            if (isa >= nSA) {
                indFull = calcSAiFromSA(G, SA, nGenome, GstrandBit, nSA - 1, Lmax, iL4);
                if (indFull == indPrev && iL4 == iL4prev) { isa = nSA; continue; }
            }
            uint64_t i1 = isa - isaStep, i2 = std::min(isa, nSA - 1);
            while (i1 + 1 < i2) {
                isa = i1 / 2 + i2 / 2 + (i1 % 2 + i2 % 2) / 2;
                if ((indFull = calcSAiFromSA(G, SA, nGenome, GstrandBit, isa, Lmax, iL4)) == indPrev && iL4 == iL4prev) i1 = isa; else i2 = isa;
            }
            if (isa == i1) { isa = i2; indFull = calcSAiFromSA(G, SA, nGenome, GstrandBit, isa, Lmax, iL4); }
        }
        for (uint32_t iL = 0; iL < Lmax; iL++)
            for (uint64_t ii = start[iL] + ind0[iL] + 1; ii < start[iL + 1]; ii++) SAi.set(ii, nSA | absentC);
    }

    // ---- the loaded-index view, then the junction inserts (Genome_genomeGenerate.cpp:337-345)
    star_index_view_t& v = idx.view;
    memset(&v, 0, sizeof(v));
    v.nGenome = nGenome; v.nSA = nSA; v.nSAbyte = nSAbyte; v.nSAi = nSAi; v.nSAibyte = nSAibyte;
    v.GstrandBit = GstrandBit; v.gSAindexNbases = Lmax; v.gSAsparseD = 1; v.gChrBinNbits = (uint32_t)P.genomeChrBinNbits; v.nChrReal = nChr;
    v.sjdbOverhang = sjdbOverhang; v.sjdbLength = sjdbLength; v.sjGstart = nGenome;
    idx.genomeDir = gDir;
    {
        const uint64_t chrBinN = idx.chrStart[nChr] / binN + 1;   // chrBinFill, Genome.cpp:209-216
        idx.chrBin.resize(chrBinN);
        for (uint64_t ii = 0, ichr = 1; ii < chrBinN; ++ii) { if (ii * binN >= idx.chrStart[ichr]) ichr++; idx.chrBin[ii] = ichr - 1; }
    }
    idx.pointView();
    if (annot) {
        HostParams Pi = P;
        Pi.sjdbInsertOutDir = gDir;
        Pi.sjdbInsertSave = "Basic";   // the index files are written below
        SjdbLoci loci;
        rc = sjdbInsertJunctions(Pi, &P.hp, idx, loci, false, "", eng, logMain, err, /*generateMode*/ true);
        if (rc) return rc;
    }
    // ---- genomeParameters.txt (genomeParametersWrite.cpp:4-46), Genome, SA, SAindex
    {
        std::ofstream gp(gDir + "genomeParameters.txt");
        gp << "### " << P.commandLineFull << "\n### GstrandBit " << GstrandBit << "\nversionGenome\t2.7.4a\ngenomeType\tFull\ngenomeFastaFiles\t";
        for (auto& f : P.genomeFastaFiles) gp << f << " ";
        gp << "\ngenomeSAindexNbases\t" << Lmax << "\ngenomeChrBinNbits\t" << P.genomeChrBinNbits << "\ngenomeSAsparseD\t1\ngenomeTransformType\tNone\ngenomeTransformVCF\t-\n";
        gp << "sjdbOverhang\t" << sjdbOverhang << "\nsjdbFileChrStartEnd\t";
        for (auto& f : P.sjdbFileChrStartEnd) gp << f << " ";
        gp << "\nsjdbGTFfile\t" << P.sjdbGTFfile << "\nsjdbGTFchrPrefix\t" << P.sjdbGTFchrPrefix << "\nsjdbGTFfeatureExon\t" << P.sjdbGTFfeatureExon
           << "\nsjdbGTFtagExonParentTranscript\t" << P.sjdbGTFtagExonParentTranscript << "\nsjdbGTFtagExonParentGene\t" << P.sjdbGTFtagExonParentGene
           << "\nsjdbInsertSave\t" << P.sjdbInsertSave << "\ngenomeFileSizes\t" << v.nGenome << " " << v.nSAbyte << "\n";
    }
    std::ofstream(gDir + "Genome", std::ios::binary).write((const char*)v.G, v.nGenome);
    std::ofstream(gDir + "SA", std::ios::binary).write((const char*)v.SA, v.nSAbyte);
    {
        std::ofstream sai(gDir + "SAindex", std::ios::binary);
        const uint64_t nb = Lmax;
        sai.write((const char*)&nb, 8);
        sai.write((const char*)idx.genomeSAindexStart.data(), 8 * (nb + 1));
        sai.write((const char*)v.SAi, v.nSAibyte);
    }
    { time_t t; time(&t); char b[100]; strftime(b, 80, "%b %d %H:%M:%S", localtime(&t)); std::cout << b << " ..... finished successfully\n" << std::flush; }
    return 0;
}

}  // namespace starhost
