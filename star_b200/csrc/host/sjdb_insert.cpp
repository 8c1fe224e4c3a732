// sjdb_insert.cpp — on-the-fly insertion of splice junctions into a loaded index.
//
// Host side of reference source/sjdbInsertJunctions.cpp:11-102 (junction lists -> inserts -> new G / SA / SAi), used by
// --sjdbFileChrStartEnd at the mapping stage and by --twopassMode Basic (twoPassRunPass1.cpp:9-96).  The two steps that touch every
// new suffix and every SA row — the suffixArraySearch1 loop and the SA rewrite of sjdbBuildIndex.cpp — run on the GPU behind
// star_gpu_sjdb_search / star_gpu_sjdb_merge_sa (include/star_b200.h, star_b200/csrc/engine/sjdb.cu); this file keeps what is small or
// strictly sequential in the reference too: list parsing, sjdbPrepare, the sort of the insertion points and the SAindex patch.
#include <sys/stat.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include "host.h"

namespace starhost {

// sjdbLoadFromStream.cpp:2-28: chr, start, end, strand ('+', '-', '.', or the 1 / 2 / 0 of SJ.out.tab); further columns are ignored
void sjdbLoadFromStream(std::istream& in, SjdbLoci& loci) {
    // A line without a 4th column leaves the reference's (uninitialised) strand variable as it was: in practice the converted strand
    // of the previous line.  Kept across the lines of one stream here, '.' for the first line.
    char s1 = 0;
    while (in.good()) {
        std::string line, chr;
        uint64_t u1 = 0, u2 = 0;
        std::getline(in, line);
        std::istringstream ls(line);
        ls >> chr >> u1 >> u2 >> s1;
        if (chr.empty()) continue;
        loci.chr.push_back(chr);
        loci.start.push_back(u1);
        loci.end.push_back(u2);
        s1 = s1 == '1' || s1 == '+' ? '+' : (s1 == '2' || s1 == '-' ? '-' : '.');
        loci.str.push_back(s1);
    }
}

namespace {

struct Packed {   // PackedArray.h:24-32, PackedArray.cpp:17-25 over a byte vector (the vector keeps >= 8 readable bytes behind the last entry)
    uint8_t* a;
    uint32_t bits;
    uint64_t mask;
    Packed(uint8_t* p, uint32_t b) : a(p), bits(b), mask(~0ULL >> (64 - b)) {}
    uint64_t get(uint64_t i) const {
        const uint64_t b = i * bits;
        uint64_t w;
        memcpy(&w, a + b / 8, 8);
        return (w >> (b % 8)) & mask;
    }
    void set(uint64_t i, uint64_t x) {
        const uint64_t b = i * bits, S = b % 8;
        uint64_t w;
        memcpy(&w, a + b / 8, 8);
        w = (w & ~(mask << S)) | (x << S);
        memcpy(a + b / 8, &w, 8);
    }
};

// funCalcSAi, SuffixArrayFuns.cpp:397-410: prefix code of iL+1 bases, negated when a base > 3 is met
int64_t calcSAi(const uint8_t* g, uint64_t iL) {
    int64_t ind = 0;
    for (uint64_t k = 0; k <= iL; k++) {
        if (g[k] > 3) return -ind;
        ind = (ind << 2) + g[k];
    }
    return ind;
}

struct Prepared {   // what sjdbPrepare leaves in the Genome object
    std::vector<uint64_t> sjdbStart, sjdbEnd, sjDstart, sjAstart;
    std::vector<uint8_t> sjdbMotif, sjdbShiftLeft, sjdbShiftRight, sjdbStrand;
};

// sjdbPrepare.cpp:5-225.  Gsj receives sjdbN inserts of sjdbLength bytes.
int sjdbPrepare(const SjdbLoci& loci, const LoadedIndex& idx, uint64_t sjdbOverhang, const std::string& outDir, Prepared& R, std::vector<uint8_t>& Gsj,
                std::ostream& logMain, std::string& err) {
    const uint8_t* G = idx.view.G;
    const uint32_t nChr = idx.view.nChrReal;
    const uint64_t nGenomeReal = idx.chrStart[nChr];
    const size_t n = loci.chr.size();
    std::vector<uint64_t> S(n), E(n);
    std::vector<uint8_t> motif(n), shL(n), shR(n);
    std::string chrOld;
    uint32_t iChr = 0;
    for (size_t ii = 0; ii < n; ii++) {
        if (chrOld != loci.chr[ii]) {
            for (iChr = 0; iChr < nChr; iChr++) if (loci.chr[ii] == idx.chrName[iChr]) break;
            if (iChr >= nChr) {
                std::ostringstream e;
                e << "EXITING because of FATAL error, the sjdb chromosome " << loci.chr[ii] << " is not found among the genomic chromosomes\n"
                  << "SOLUTION: fix your file(s) --sjdbFileChrStartEnd or --sjdbGTFfile, offending junction:" << loci.chr[ii] << "\t" << loci.start[ii] << "\t" << loci.end[ii] << "\n";
                err = e.str();
                return STAR_EXIT_INPUT_FILES;
            }
            chrOld = loci.chr[ii];
        }
        uint64_t s = loci.start[ii] + idx.chrStart[iChr] - 1, e = loci.end[ii] + idx.chrStart[iChr] - 1;   // 1-based intron loci
        const uint8_t d0 = G[s], d1 = G[s + 1], a0 = G[e - 1], a1 = G[e];
        if (d0 == 2 && d1 == 3 && a0 == 0 && a1 == 2) motif[ii] = 1;        // GT/AG
        else if (d0 == 1 && d1 == 3 && a0 == 0 && a1 == 1) motif[ii] = 2;   // CT/AC
        else if (d0 == 2 && d1 == 1 && a0 == 0 && a1 == 2) motif[ii] = 3;   // GC/AG
        else if (d0 == 1 && d1 == 3 && a0 == 2 && a1 == 1) motif[ii] = 4;   // CT/GC
        else if (d0 == 0 && d1 == 3 && a0 == 0 && a1 == 1) motif[ii] = 5;   // AT/AC
        else if (d0 == 2 && d1 == 3 && a0 == 0 && a1 == 3) motif[ii] = 6;   // GT/AT
        else motif[ii] = 0;
        uint64_t jjL = 0, jjR = 0;   // repeat around the junction
        while (jjL <= s - 1 && G[s - 1 - jjL] == G[e - jjL] && G[s - 1 - jjL] < 4 && jjL < 255) jjL++;
        while (s + jjR < nGenomeReal && G[s + jjR] == G[e + 1 + jjR] && G[s + jjR] < 4 && jjR < 255) jjR++;
        shL[ii] = (uint8_t)jjL; shR[ii] = (uint8_t)jjR;
        if (jjR == 255 || jjL == 255)
            logMain << "WARNING: long repeat for junction # " << ii + 1 << " : " << loci.chr[ii] << " " << s - idx.chrStart[iChr] + 1 << " " << e - idx.chrStart[iChr] + 1
                    << "; left shift = " << (int)shL[ii] << "; right shift = " << (int)shR[ii] << "\n";
        S[ii] = s - jjL; E[ii] = e - jjL;
    }
    // first sort: strand classes apart, then (start, end) of the left-flushed junction; duplicates resolved by priority / motif / shift
    struct Key { uint64_t a, b, i; };
    auto byAB = [](const Key& x, const Key& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; };   // funCompareUint2
    std::vector<Key> srt(n);
    for (size_t ii = 0; ii < n; ii++) {
        const uint64_t shift1 = loci.str[ii] == '+' ? 0 : (loci.str[ii] == '-' ? nGenomeReal : 2 * nGenomeReal);
        srt[ii] = Key{S[ii] + shift1, E[ii] + shift1, ii};
    }
    std::stable_sort(srt.begin(), srt.end(), byAB);
    std::vector<uint64_t> I;
    I.reserve(n);
    for (size_t ii = 0; ii < n; ii++) {
        const uint64_t isj = srt[ii].i;
        if (I.empty()) { I.push_back(isj); continue; }
        const uint64_t isj0 = I.back();
        if (S[isj] != S[isj0] || E[isj] != E[isj0]) I.push_back(isj);
        else if (loci.priority[isj] < loci.priority[isj0]) {}
        else if (loci.priority[isj] > loci.priority[isj0]) I.back() = isj;
        else if ((motif[isj] > 0 && motif[isj0] == 0) || (((motif[isj] > 0) == (motif[isj0] > 0)) && shL[isj] < shL[isj0])) I.back() = isj;
    }
    // second sort: canonical junctions back at their true loci
    const size_t nsj = I.size();
    srt.resize(nsj);
    for (size_t ii = 0; ii < nsj; ii++) {
        const uint64_t back = motif[I[ii]] == 0 ? 0 : shL[I[ii]];
        srt[ii] = Key{S[I[ii]] + back, E[I[ii]] + back, I[ii]};
    }
    std::stable_sort(srt.begin(), srt.end(), byAB);
    R = Prepared();
    auto& st = R.sjdbStart; auto& en = R.sjdbEnd; auto& mo = R.sjdbMotif; auto& sl = R.sjdbShiftLeft; auto& sr = R.sjdbShiftRight; auto& sd = R.sjdbStrand;
    for (size_t ii = 0; ii < nsj; ii++) {
        const uint64_t isj = srt[ii].i;
        if (!st.empty() && st.back() == srt[ii].a && en.back() == srt[ii].b) {   // the same loci on opposite strands
            const uint64_t isj0 = srt[ii - 1].i;
            if (loci.priority[isj] < loci.priority[isj0]) continue;
            else if (loci.priority[isj] > loci.priority[isj0]) {}                          // replace
            else if (sd.back() > 0 && loci.str[isj] == '.') continue;
            else if (sd.back() == 0 && loci.str[isj] != '.') {}                            // replace
            else if (mo.back() == 0 && motif[isj] == 0) { sd.back() = 0; continue; }       // both non-canonical: strand undefined
            else if ((mo.back() > 0 && motif[isj] == 0) || (mo.back() % 2 == (2 - sd.back()))) continue;
            st.pop_back(); en.pop_back(); mo.pop_back(); sl.pop_back(); sr.pop_back(); sd.pop_back();
        }
        st.push_back(srt[ii].a); en.push_back(srt[ii].b); mo.push_back(motif[isj]); sl.push_back(shL[isj]); sr.push_back(shR[isj]);
        if (loci.str[isj] == '+') sd.push_back(1);
        else if (loci.str[isj] == '-') sd.push_back(2);
        else sd.push_back(motif[isj] == 0 ? 0 : 2 - motif[isj] % 2);
    }
    const uint64_t sjdbN = st.size(), sjdbLength = 2 * sjdbOverhang + 1;
    R.sjDstart.resize(sjdbN); R.sjAstart.resize(sjdbN);
    Gsj.assign(2 * sjdbLength * sjdbN + 1, 5);
    std::ofstream sjdbInfo(outDir + "/sjdbInfo.txt"), sjdbList(outDir + "/sjdbList.out.tab");
    const char strandChar[3] = {'.', '+', '-'};
    sjdbInfo << sjdbN << "\t" << sjdbOverhang << "\n";
    for (uint64_t ii = 0; ii < sjdbN; ii++) {
        R.sjDstart[ii] = st[ii] - sjdbOverhang;
        R.sjAstart[ii] = en[ii] + 1;
        if (mo[ii] == 0) { R.sjDstart[ii] += sl[ii]; R.sjAstart[ii] += sl[ii]; }   // non-canonical: true coordinates
        memcpy(Gsj.data() + ii * sjdbLength, G + R.sjDstart[ii], sjdbOverhang);
        memcpy(Gsj.data() + ii * sjdbLength + sjdbOverhang, G + R.sjAstart[ii], sjdbOverhang);
        Gsj[(ii + 1) * sjdbLength - 1] = 5;   // GENOME_spacingChar between the inserts
        sjdbInfo << st[ii] << "\t" << en[ii] << "\t" << (int)mo[ii] << "\t" << (int)sl[ii] << "\t" << (int)sr[ii] << "\t" << (int)sd[ii] << "\n";
        const uint64_t chr1 = idx.chrBin[st[ii] >> idx.view.gChrBinNbits];
        const uint64_t back = mo[ii] > 0 ? 0 : sl[ii];
        sjdbList << idx.chrName[chr1] << "\t" << st[ii] - idx.chrStart[chr1] + 1 + back << "\t" << en[ii] - idx.chrStart[chr1] + 1 + back << "\t" << strandChar[sd[ii]] << "\n";
    }
    return 0;
}

// binarySearch2.cpp: index of (x,y) in the (X,Y) list sorted by X then Y, or -1
int64_t findJunction(uint64_t x, uint64_t y, const std::vector<uint64_t>& X, const std::vector<uint64_t>& Y) {
    auto it = std::lower_bound(X.begin(), X.end(), x);
    for (size_t i = it - X.begin(); i < X.size() && X[i] == x; i++) if (Y[i] == y) return (int64_t)i;
    return -1;
}

bool copyFile(const std::string& a, const std::string& b) {
    std::ifstream in(a, std::ios::binary);
    std::ofstream out(b, std::ios::binary);
    if (!in.good() || !out.good()) return false;
    out << in.rdbuf();
    return true;
}

}  // namespace

// The junction part of GTF::GTF + GTF::transcriptGeneSJ (GTF.cpp:7-200, GTF_transcriptGeneSJ.cpp:23-183): exon lines -> exons per
// transcript -> introns between consecutive exons, collapsed by (start, end, strand); writes sjdbList.fromGTF.out.tab and the transcript /
// exon / gene tables of the reference (exonInfo.tab, transcriptInfo.tab, geneInfo.tab, exonGeTrInfo.tab; they serve --quantMode).
static int sjdbLoadFromGTF(const HostParams& P, const LoadedIndex& idx, const std::string& outDir, SjdbLoci& loci, std::ostream& logMain, std::string& err) {
    std::ifstream in(P.sjdbGTFfile);
    if (in.fail()) { err = "FATAL error, could not open file pGe.sjdbGTFfile=" + P.sjdbGTFfile + "\n"; return STAR_EXIT_INPUT_FILES; }
    std::map<std::string, uint32_t> chrIndex;
    for (uint32_t i = 0; i < idx.view.nChrReal; i++) chrIndex[idx.chrName[i]] = i;
    std::map<std::string, uint64_t> trNumber, geNumber;
    std::vector<uint8_t> trStrand;
    std::vector<std::string> transcriptID, geneID;
    std::vector<std::array<std::string, 2>> geneAttr;
    struct Ex { uint64_t t, s, e, g; };
    std::vector<Ex> ex;
    uint64_t nExonLines = 0;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string chr1, f2, feature;
        ls >> chr1 >> f2 >> feature;
        if (chr1.substr(0, 1) == "#" || feature != P.sjdbGTFfeatureExon) continue;
        nExonLines++;
        if (P.sjdbGTFchrPrefix != "-") chr1 = P.sjdbGTFchrPrefix + chr1;
        auto ci = chrIndex.find(chr1);
        if (ci == chrIndex.end()) {
            logMain << "WARNING: while processing sjdbGTFfile=" << P.sjdbGTFfile << ": chromosome '" << chr1 << "' not found in Genome fasta files for line:\n" << line << "\n";
            continue;
        }
        uint64_t ex1 = 0, ex2 = 0;
        char str1 = '.';
        ls >> ex1 >> ex2 >> f2 >> str1 >> f2;
        if (ex2 > idx.chrLength[ci->second]) {
            logMain << "WARNING: while processing sjdbGTFfile=" << P.sjdbGTFfile << ", line:\n" << line << "\n exon end = " << ex2 << " is larger than the chromosome " << chr1
                    << " length = " << idx.chrLength[ci->second] << " , will skip this exon\n";
            continue;
        }
        std::string attrs;
        std::getline(ls, attrs);
        for (char& c : attrs) if (c == ';' || c == '=' || c == '\t' || c == '"') c = ' ';
        auto attr = [&](const std::string& name) {
            std::string v;
            size_t pos = attrs.find(" " + name + " ");
            if (pos != std::string::npos) pos = attrs.find_first_not_of(" ", pos + name.size() + 1);
            if (pos != std::string::npos) v = attrs.substr(pos, attrs.find_first_of(" ", pos) - pos);
            return v;
        };
        std::string trID = attr(P.sjdbGTFtagExonParentTranscript), gID = attr(P.sjdbGTFtagExonParentGene), gName, gType;
        for (const std::string& nm : P.sjdbGTFtagExonParentGeneName) { std::string v = attr(nm); if (!v.empty()) gName = v; }
        for (const std::string& nm : P.sjdbGTFtagExonParentGeneType) { std::string v = attr(nm); if (!v.empty()) gType = v; }
        if (trID.empty()) {
            logMain << "WARNING: while processing pGe.sjdbGTFfile=" << P.sjdbGTFfile << ": no transcript_id for line:\n" << line << "\n";
            trID = "tr_" + chr1 + "_" + std::to_string(ex1) + "_" + std::to_string(ex2) + "_" + std::to_string(ex.size());
        }
        if (gID.empty()) {
            logMain << "WARNING: while processing pGe.sjdbGTFfile=" << P.sjdbGTFfile << ": no gene_id for line:\n" << line << "\n";
            gID = "MissingGeneID";
        }
        if (gName.empty()) gName = gID;
        if (gType.empty()) gType = "MissingGeneType";
        auto ti = trNumber.insert({trID, trNumber.size()});
        if (ti.second) { trStrand.push_back(str1 == '+' ? 1 : (str1 == '-' ? 2 : 0)); transcriptID.push_back(trID); }
        auto gi = geNumber.insert({gID, geNumber.size()});
        if (gi.second) { geneID.push_back(gID); geneAttr.push_back({gName, gType}); }
        ex.push_back(Ex{ti.first->second, ex1 + idx.chrStart[ci->second] - 1, ex2 + idx.chrStart[ci->second] - 1, gi.first->second});
    }
    if (nExonLines == 0) {
        err = "Fatal INPUT FILE error, no exon lines in the GTF file: " + P.sjdbGTFfile + "\nSolution: check the formatting of the GTF file, it must contain some lines with exon in the 3rd column.\n          Make sure the GTF file is unzipped.\n          If exons are marked with a different word, use --sjdbGTFfeatureExon .\n";
        return STAR_EXIT_INPUT_FILES;
    }
    if (ex.empty()) {
        err = "Fatal INPUT FILE error, no valid exon lines in the GTF file: " + P.sjdbGTFfile + "\nSolution: check the formatting of the GTF file. One likely cause is the difference in chromosome naming between GTF and FASTA file.\n";
        return STAR_EXIT_INPUT_FILES;
    }
    std::stable_sort(ex.begin(), ex.end(), [](const Ex& a, const Ex& b) { return a.t != b.t ? a.t < b.t : a.s < b.s; });   // funCompareUint2 on (transcript, start)
    {   // the annotation tables next to the index (GTF_transcriptGeneSJ.cpp:32-114): exons by locus, genes, transcripts with their exons
        const uint64_t exonN = ex.size();
        std::vector<std::array<uint64_t, 5>> exge(exonN);
        for (uint64_t i = 0; i < exonN; i++) exge[i] = {ex[i].s, ex[i].e, (uint64_t)trStrand[ex[i].t], ex[i].g, ex[i].t};
        std::sort(exge.begin(), exge.end());
        std::ofstream exgeOut(outDir + "/exonGeTrInfo.tab");
        exgeOut << exonN << "\n";
        for (auto& r : exge) exgeOut << r[0] << "\t" << r[1] << "\t" << r[2] << "\t" << r[3] << "\t" << r[4] << "\n";
        std::ofstream geOut(outDir + "/geneInfo.tab");
        geOut << geneID.size() << "\n";
        for (size_t ig = 0; ig < geneID.size(); ig++) geOut << geneID[ig] << "\t" << geneAttr[ig][0] << "\t" << geneAttr[ig][1] << "\n";
        // rows (transcript start, transcript end, transcript, exon start, exon end, gene) sorted by the first five
        std::vector<std::array<uint64_t, 6>> extr(exonN);
        uint64_t trex1 = 0;
        for (uint64_t iex = 0; iex <= exonN; iex++) {
            if (iex == exonN || ex[iex].t != ex[trex1].t) {
                for (uint64_t k = trex1; k < iex; k++) extr[k][1] = ex[iex - 1].e;   // (the end of the transcript's LAST exon by start)
                if (iex == exonN) break;
                trex1 = iex;
            }
            extr[iex][0] = ex[trex1].s; extr[iex][2] = ex[iex].t; extr[iex][3] = ex[iex].s; extr[iex][4] = ex[iex].e; extr[iex][5] = ex[iex].g;
        }
        std::stable_sort(extr.begin(), extr.end(), [](const std::array<uint64_t, 6>& a, const std::array<uint64_t, 6>& b) {
            for (int k = 0; k < 5; k++) if (a[k] != b[k]) return a[k] < b[k];
            return false;
        });
        std::ofstream trOut(outDir + "/transcriptInfo.tab"), exOut(outDir + "/exonInfo.tab");
        trOut << transcriptID.size() << "\n";
        exOut << exonN << "\n";
        uint64_t trid = extr[0][2], trex = 0, trstart = extr[0][0], trend = extr[0][1], exlen = 0;
        for (uint64_t iex = 0; iex <= exonN; iex++) {
            if (iex == exonN || extr[iex][2] != trid) {
                trOut << transcriptID.at(trid) << "\t" << extr[iex - 1][0] << "\t" << extr[iex - 1][1] << "\t" << trend << "\t" << (uint64_t)trStrand[trid] << "\t" << iex - trex << "\t"
                      << trex << "\t" << extr[iex - 1][5] << "\n";
                if (iex == exonN) break;
                trid = extr[iex][2]; trstart = extr[iex][0]; trex = iex;
                trend = std::max(trend, extr[iex - 1][1]);
                exlen = 0;
            }
            exOut << extr[iex][3] - trstart << "\t" << extr[iex][4] - trstart << "\t" << exlen << "\n";
            exlen += extr[iex][4] - extr[iex][3] + 1;
        }
    }
    struct Sj { uint64_t s, e, str, g; };
    std::vector<Sj> sj;
    uint64_t trCur = ex[0].t;
    for (size_t i = 1; i < ex.size(); i++) {
        if (trCur != ex[i].t) { trCur = ex[i].t; continue; }
        if (ex[i].s <= ex[i - 1].e + 1) continue;   // touching or overlapping exons: no intron
        sj.push_back(Sj{ex[i - 1].e + 1, ex[i].s - 1, trStrand[trCur], ex[i].g + 1});
    }
    std::stable_sort(sj.begin(), sj.end(), [](const Sj& a, const Sj& b) { return a.s != b.s ? a.s < b.s : a.e < b.e; });
    const char strandChar[3] = {'.', '+', '-'};
    const size_t n0 = loci.chr.size();
    std::vector<std::set<uint64_t>> genes;
    for (size_t i = 0; i < sj.size(); i++) {
        if (i == 0 || sj[i].s != sj[i - 1].s || sj[i].e != sj[i - 1].e || sj[i].str != sj[i - 1].str) {
            const uint64_t chr1 = idx.chrBin[sj[i].s >> idx.view.gChrBinNbits];
            loci.chr.push_back(idx.chrName[chr1]);
            loci.start.push_back(sj[i].s + 1 - idx.chrStart[chr1]);
            loci.end.push_back(sj[i].e + 1 - idx.chrStart[chr1]);
            loci.str.push_back(strandChar[sj[i].str]);
            genes.push_back({sj[i].g});
        } else genes.back().insert(sj[i].g);
    }
    std::ofstream list(outDir + "/sjdbList.fromGTF.out.tab");
    for (size_t i = n0; i < loci.chr.size(); i++) {
        list << loci.chr[i] << "\t" << loci.start[i] << "\t" << loci.end[i] << "\t" << loci.str[i];
        bool first = true;
        for (uint64_t g : genes[i - n0]) { list << (first ? "\t" : ",") << g; first = false; }
        list << "\n";
    }
    loci.priority.resize(loci.chr.size(), 20);
    logMain << "Processing pGe.sjdbGTFfile=" << P.sjdbGTFfile << ", found:\n\t\t" << trNumber.size() << " transcripts\n\t\t" << ex.size() << " exons (non-collapsed)\n\t\t"
            << loci.chr.size() - n0 << " collapsed junctions\nTotal junctions: " << loci.chr.size() << "\n";
    return 0;
}

int sjdbInsertJunctions(const HostParams& P, star_params_t* hp, LoadedIndex& idx, SjdbLoci& loci, bool pass2, const std::string& pass1sjFile,
                        const star_engine_vtbl_t* eng, std::ostream& logMain, std::string& err, bool generateMode) {
    star_index_view_t& v = idx.view;
    const std::string& outDir = P.sjdbInsertOutDir;
    if (v.sjdbN > 0 && loci.chr.empty()) {   // junctions of the generated genome (only if they were not loaded before)
        std::ifstream in(idx.genomeDir + "/sjdbList.out.tab");
        if (in.fail()) { err = "EXITING because of fatal INPUT error: could not open input file " + idx.genomeDir + "/sjdbList.out.tab\nSOLUTION: re-generate the genome in pGe.gDir=" + idx.genomeDir + "\n"; return STAR_EXIT_INPUT_FILES; }
        sjdbLoadFromStream(in, loci);
        loci.priority.resize(loci.chr.size(), 30);
        logMain << "   Loaded database junctions from the generated genome " << idx.genomeDir << "/sjdbList.out.tab: " << loci.chr.size() << " total junctions\n\n";
    }
    if (pass2) {   // the junctions found in the 1st pass
        std::ifstream in(pass1sjFile);
        if (in.fail()) { err = "FATAL INPUT error, could not open input file with junctions from the 1st pass=" + pass1sjFile + "\n"; return STAR_EXIT_INPUT_FILES; }
        sjdbLoadFromStream(in, loci);
        loci.priority.resize(loci.chr.size(), 0);
        logMain << "   Loaded database junctions from the 1st pass file: " << pass1sjFile << ": " << loci.chr.size() << " total junctions\n\n";
    } else {
        if (generateMode && P.sjdbGTFfile != "-") {   // at genome generation the GTF junctions come first (Genome_genomeGenerate.cpp:160-163)
            int rcg = sjdbLoadFromGTF(P, idx, outDir, loci, logMain, err);
            if (rcg) return rcg;
        }
        if (P.sjdbFileChrStartEnd[0] != "-")   // sjdbLoadFromFiles.cpp:6-26
            for (const std::string& fn : P.sjdbFileChrStartEnd) {
                std::ifstream in(fn);
                if (in.fail()) { err = "FATAL INPUT error, could not open input file pGe.sjdbFileChrStartEnd=" + fn + "\n"; return STAR_EXIT_INPUT_FILES; }
                sjdbLoadFromStream(in, loci);
                loci.priority.resize(loci.chr.size(), 10);
                logMain << "   Loaded database junctions from the pGe.sjdbFileChrStartEnd file(s), total number of junctions:" << loci.chr.size() << "\n\n";
            }
        if (!generateMode && P.sjdbGTFfile != "-") {
            int rcg = sjdbLoadFromGTF(P, idx, outDir, loci, logMain, err);
            if (rcg) return rcg;
        }
    }
    const uint64_t sjdbOverhang = v.sjdbOverhang, sjdbLength = v.sjdbLength;
    const uint64_t nGenomeReal = idx.chrStart[v.nChrReal];
    Prepared R;
    std::vector<uint8_t> Gsj;
    int rc = sjdbPrepare(loci, idx, sjdbOverhang, outDir, R, Gsj, logMain, err);
    if (rc) return rc;
    logMain << "   Finished preparing junctions" << std::endl;
    const uint64_t sjdbN = R.sjdbStart.size();
    if (sjdbN > P.limitSjdbInsertNsj) {
        std::ostringstream e;
        e << "Fatal LIMIT error: the number of junctions to be inserted on the fly =" << sjdbN << " is larger than the limitSjdbInsertNsj=" << P.limitSjdbInsertNsj << "\n";
        e << "Fatal LIMIT error: the number of junctions to be inserted on the fly =" << sjdbN << " is larger than the limitSjdbInsertNsj=" << P.limitSjdbInsertNsj << "\n";
        e << "SOLUTION: re-run with at least --limitSjdbInsertNsj " << sjdbN << "\n";
        err = e.str();
        return STAR_EXIT_INPUT_FILES;
    }

    // ---- sjdbBuildIndex.cpp:16-333
    if (sjdbN > 0) {
        logMain << " ..... inserting junctions into the genome indices" << std::endl;
        const uint64_t nGsj = sjdbLength * sjdbN;
        for (uint64_t ii = 0; ii < nGsj; ii++) Gsj[2 * nGsj - 1 - ii] = Gsj[ii] < 4 ? 3 - Gsj[ii] : Gsj[ii];   // reverse complement of the inserts
        Gsj[2 * nGsj] = 5;
        // junctions that are in the index already keep their rows: no suffixes for them, only a new position
        const uint64_t sjdbNold = v.sjdbN;
        std::vector<uint32_t> oldSJind(std::max<uint64_t>(1, sjdbNold), 0);
        std::vector<uint8_t> skipSeq(2 * sjdbN, 0);
        uint64_t sjNew = 0;
        for (uint64_t isj = 0; isj < 2 * sjdbN; isj++) {
            const uint64_t isj1 = isj < sjdbN ? isj : 2 * sjdbN - 1 - isj;
            const int64_t old = sjdbNold == 0 ? -1 : findJunction(R.sjdbStart[isj1], R.sjdbEnd[isj1], idx.sjdbStart, idx.sjdbEnd);
            if (old < 0) ++sjNew;
            else { oldSJind[old] = (uint32_t)isj1; skipSeq[isj] = 1; }
        }
        sjNew /= 2;
        const uint64_t nSuf = 2 * sjdbN * sjdbLength;
        std::vector<uint64_t> ind(2 * (nSuf + 1));
        void* h = nullptr;
        rc = eng->sjdb_open(&h, P.gpuDevice, &v);
        if (!rc) rc = eng->sjdb_search(h, Gsj.data(), sjdbN, sjdbLength, skipSeq.data(), ind.data());
        if (rc) { if (h) eng->sjdb_close(h); err = std::string("EXITING because of FATAL ERROR: junction insertion failed: ") + eng->last_error() + "\n"; return rc; }
        logMain << "   Finished SA search: number of new junctions=" << sjNew << ", old junctions=" << sjdbN - sjNew << std::endl;
        struct Ins { uint64_t row, off; };
        static_assert(sizeof(Ins) == 16, "pairs of 64-bit words");
        Ins* ia = reinterpret_cast<Ins*>(ind.data());
        uint64_t nInd = 0;
        for (uint64_t ii = 0; ii < nSuf; ii++) if (ia[ii].row != ~0ULL) ia[nInd++] = ia[ii];
        const uint8_t* gs = Gsj.data();
        std::sort(ia, ia + nInd, [gs](const Ins& x, const Ins& y) {   // funCompareUintAndSuffixes.cpp:6-43: row, then suffix text up to a common 5, then offset
            if (x.row != y.row) return x.row < y.row;
            const uint8_t *ga = gs + x.off, *gb = gs + y.off;
            for (uint64_t ig = 0;; ig++) {
                if (ga[ig] != gb[ig]) return ga[ig] < gb[ig];
                if (ga[ig] == 5) return x.off < y.off;
            }
        });
        logMain << "   Finished sorting SA indicesL nInd=" << nInd << std::endl;
        ia[nInd].row = (uint64_t)-999; ia[nInd].off = (uint64_t)-999;   // end marker read by the SAindex loops below

        const uint64_t nGenomeNew = nGenomeReal + nGsj, nSAnew = v.nSA + nInd;
        uint32_t GstrandBit1 = (uint32_t)std::floor(std::log((double)nGenomeNew) / std::log(2.0)) + 1;
        if (GstrandBit1 < 32) GstrandBit1 = 32;
        logMain << "Genome size with junctions=" << nGenomeNew << "  " << nGenomeReal << "   " << nGsj << "\nGstrandBit1=" << GstrandBit1 << "   GstrandBit=" << v.GstrandBit << "\n";
        if (GstrandBit1 > v.GstrandBit) {
            eng->sjdb_close(h);
            err = "EXITING because of FATAL ERROR: cannot insert junctions on the fly because of strand GstrandBit problem\nSOLUTION: please contact STAR author at https://groups.google.com/forum/#!forum/rna-star\n";
            return STAR_EXIT_GENOME_FILES;
        }
        const uint64_t nSAnewByte = (nSAnew - 1) * (v.GstrandBit + 1) / 8 + 8;
        std::vector<uint8_t> SAnew(nSAnewByte + 16, 0);
        rc = eng->sjdb_merge_sa(h, ind.data(), nInd, nGsj, sjNew * sjdbLength, sjdbLength, oldSJind.data(), SAnew.data(), nSAnewByte);
        eng->sjdb_close(h);
        if (rc) { err = std::string("EXITING because of FATAL ERROR: junction insertion failed: ") + eng->last_error() + "\n"; return rc; }
        logMain << "   Finished inserting junction indices" << std::endl;

        // SAindex: every prefix length, one sweep with a running count of inserted rows (sjdbBuildIndex.cpp:219-268)
        Packed SAi(idx.SAistore.data(), v.GstrandBit + 3);
        const uint64_t absentC = 1ULL << (v.GstrandBit + 2), nC = 1ULL << (v.GstrandBit + 1);
        const uint64_t* start = idx.genomeSAindexStart.data();
        for (uint64_t iL = 0; iL < v.gSAindexNbases; iL++) {
            uint64_t iSJ = 0;
            uint64_t ind0 = start[iL] - 1;   // last prefix that was present
            for (uint64_t ii = start[iL]; ii < start[iL + 1]; ii++) {
                const uint64_t iSA1 = SAi.get(ii);
                const uint64_t iSA2 = iSA1 & ~nC & ~absentC;
                if (iSJ < nInd && (iSA1 & absentC) > 0) {   // prefix absent from the old genome: present now if an insert carries it
                    const uint64_t iSJ1 = iSJ;
                    int64_t ind1 = calcSAi(gs + ia[iSJ].off, iL);
                    while (ind1 < (int64_t)(ii - start[iL]) && ia[iSJ].row - 1 < iSA2) {
                        ++iSJ;
                        ind1 = calcSAi(gs + ia[iSJ].off, iL);
                    }
                    if (ind1 == (int64_t)(ii - start[iL])) {
                        SAi.set(ii, ia[iSJ].row - 1 + iSJ + 1);
                        for (uint64_t ii0 = ind0 + 1; ii0 < ii; ii0++) SAi.set(ii0, (ia[iSJ].row - 1 + iSJ + 1) | absentC);
                        ++iSJ;
                        ind0 = ii;
                    } else iSJ = iSJ1;
                } else {   // present before: shift by the inserted rows in front of it
                    while (iSJ < nInd && ia[iSJ].row - 1 + 1 < iSA2) ++iSJ;
                    while (iSJ < nInd && ia[iSJ].row - 1 + 1 == iSA2) {   // insertions right at this row: those with a smaller prefix go in front
                        if (calcSAi(gs + ia[iSJ].off, iL) >= (int64_t)(ii - start[iL])) break;
                        ++iSJ;
                    }
                    SAi.set(ii, iSA1 + iSJ);
                    for (uint64_t ii0 = ind0 + 1; ii0 < ii; ii0++) SAi.set(ii0, (iSA2 + iSJ) | absentC);
                    ind0 = ii;
                }
            }
        }
        for (uint64_t isj = 0; isj < nInd; isj++) {   // inserts that meet a base > 3 within the prefix: flag "contains N" (:270-293)
            int64_t ind1 = 0;
            for (uint64_t iL = 0; iL < v.gSAindexNbases; iL++) {
                const uint64_t g = gs[ia[isj].off + iL];
                ind1 <<= 2;
                if (g > 3) {
                    for (uint64_t iL1 = iL; iL1 < v.gSAindexNbases; iL1++) {
                        ind1 += 3;
                        int64_t ind2 = (int64_t)start[iL1] + ind1;
                        for (; ind2 >= 0; ind2--) if ((SAi.get(ind2) & absentC) == 0) break;
                        SAi.set(ind2, SAi.get(ind2) | nC);
                        ind1 <<= 2;
                    }
                    break;
                }
                ind1 += g;
            }
        }
        logMain << "   Finished SAi" << std::endl;

        // the new genome: real chromosomes, then the inserts (:296-300)
        std::vector<uint8_t> Gnew(256 + nGenomeNew + 256, 5);
        memcpy(Gnew.data() + 256, idx.Gstore.data() + 256, nGenomeReal);
        memcpy(Gnew.data() + 256 + nGenomeReal, Gsj.data(), nGsj);
        idx.Gstore.swap(Gnew);
        idx.SAstore.swap(SAnew);
        v.nGenome = nGenomeNew; v.nSA = nSAnew; v.nSAbyte = nSAnewByte;
        v.sjGstart = nGenomeReal;
    }
    idx.sjdbStart = R.sjdbStart; idx.sjdbEnd = R.sjdbEnd; idx.sjDstart = R.sjDstart; idx.sjAstart = R.sjAstart;
    idx.sjdbMotif = R.sjdbMotif; idx.sjdbShiftLeft = R.sjdbShiftLeft; idx.sjdbShiftRight = R.sjdbShiftRight; idx.sjdbStrand = R.sjdbStrand;
    v.sjdbN = sjdbN;
    idx.pointView();
    logMain << " ..... finished inserting junctions into genome" << std::endl;

    if (P.sjdbInsertSave == "All") {   // the whole index, loadable by STAR and by loadIndex (sjdbInsertJunctions.cpp:69-97)
        if (idx.genomeDir != outDir)
            for (const char* f : {"/chrName.txt", "/chrStart.txt", "/chrNameLength.txt", "/chrLength.txt"}) copyFile(idx.genomeDir + f, outDir + f);
        {
            std::ifstream in(idx.genomeDir + "/genomeParameters.txt");   // the lines of the loaded index with the sizes and junction settings replaced
            std::ofstream out(outDir + "/genomeParameters.txt");
            std::string line;
            while (std::getline(in, line)) {
                std::istringstream ls(line);
                std::string w1;
                ls >> w1;
                if (w1 == "genomeFileSizes") out << "genomeFileSizes\t" << v.nGenome << " " << v.nSAbyte << "\n";
                else if (w1 == "sjdbOverhang") out << "sjdbOverhang\t" << sjdbOverhang << "\n";
                else if (w1 == "sjdbInsertSave") out << "sjdbInsertSave\t" << P.sjdbInsertSave << "\n";
                else if (w1 == "sjdbFileChrStartEnd") { out << "sjdbFileChrStartEnd\t"; for (auto& f : P.sjdbFileChrStartEnd) out << f << " "; out << "\n"; }
                else out << line << "\n";
            }
        }
        std::ofstream(outDir + "/Genome", std::ios::binary).write((const char*)v.G, v.nGenome);
        std::ofstream(outDir + "/SA", std::ios::binary).write((const char*)v.SA, v.nSAbyte);
        std::ofstream sai(outDir + "/SAindex", std::ios::binary);
        const uint64_t nb = v.gSAindexNbases;
        sai.write((const char*)&nb, 8);
        sai.write((const char*)idx.genomeSAindexStart.data(), 8 * (nb + 1));
        sai.write((const char*)v.SAi, v.nSAibyte);
    }
    hp->winBinN = v.nGenome / (1ULL << hp->winBinNbits) + 1;
    return 0;
}

}  // namespace starhost
