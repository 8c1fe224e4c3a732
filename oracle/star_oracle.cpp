/*
 * star_oracle.cpp — CPU restatement of STAR's per-read alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may link, load or execute this file; the product (star_b200/) never does and has no CPU fallback.
 *
 * This is a from-scratch restatement (flat PODs, no STL containers inside the per-read state) of
 * the algorithm in alexdobin/STAR 2.7.11b; every function cites the reference file:line it follows
 * (paths relative to /root/reference/source).  Integer types mirror the reference's
 * (`uint` = unsigned long long, `int` = 32 bit) because several comparisons rely on wrap-around.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py runs the unmodified reference
 * (oracle/_ref/STAR, built by oracle/Makefile.ref) and this restatement on the same seeded inputs
 * and requires byte-identical SAM records, SJ.out.tab and Log.final.out counters; the committed
 * fixtures in tests/golden/ were produced by the reference binary (tests/golden/make_golden.py).
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/star_b200.h"
#include "star_oracle.h"

#define uint unsigned long long   // exactly the reference's definition, IncludeDefine.h:50
typedef int intScore;

namespace {

enum { PC_rStart = 0, PC_Length, PC_Str, PC_Dir, PC_Nrep, PC_SAstart, PC_SAend, PC_iFrag, PC_SIZE };  // IncludeDefine.h:181-189
enum { WC_Str = 0, WC_Chr, WC_gStart, WC_gEnd, WC_SIZE };                                          // :191-195
enum { WA_Length = 0, WA_rStart, WA_gStart, WA_Nrep, WA_Anchor, WA_iFrag, WA_sjA, WA_SIZE };       // :197-204
enum { EX_R = 0, EX_G, EX_L, EX_iFrag, EX_sjA, EX_SIZE };                                          // :206-211
const int MAX_N_EXONS = STAR_MAX_N_EXONS;
const uint MAX_SJ_REPEAT_SEARCH = 255;   // IncludeDefine.h:176
const int scoreMatch = 1;                // IncludeDefine.h:70
const uint DEF_readSeqLengthMax = STAR_READ_SEQ_LENGTH_MAX;
const char MARK_FRAG_SPACER_BASE = STAR_MARK_FRAG_SPACER_BASE;
const int TOO_MANY_WINDOWS = 101;        // EXIT_createExtendWindowsWithAlign_TOO_MANY_WINDOWS
const unsigned short uintWinBinMax = 65535;

// Transcript.h:10-81 without the std:: members that the path never touches
struct Tr {
    uint exons[MAX_N_EXONS][EX_SIZE];
    uint shiftSJ[MAX_N_EXONS][2];
    int canonSJ[MAX_N_EXONS];
    uint8_t sjAnnot[MAX_N_EXONS];
    uint8_t sjStr[MAX_N_EXONS];
    uint intronMotifs[3];
    uint8_t sjMotifStrand;
    bool sjYes;
    uint nExons;
    int iFrag;
    uint rStart, roStart, rLength, gStart, gLength, cStart;
    uint Chr, Str, roStr;
    bool primaryFlag;
    uint nMatch, nMM, mappedLength, extendL;
    intScore maxScore;
    uint nGap, lGap, nDel, nIns, lDel, lIns;
    uint nUnique, nAnchor;

    void reset() {  // Transcript.cpp:8-26
        extendL = 0; primaryFlag = false;
        rStart = 0; roStart = 0; rLength = 0; gStart = 0; gLength = 0;
        maxScore = 0; nMatch = 0; nMM = 0;
        nGap = 0; lGap = 0; lDel = 0; lIns = 0; nDel = 0; nIns = 0;
        nUnique = nAnchor = 0;
    }
    void add(const Tr* t) {  // Transcript.cpp:28-36
        maxScore += t->maxScore; nMatch += t->nMatch; nMM += t->nMM;
        nGap += t->nGap; lGap += t->lGap; lDel += t->lDel; nDel += t->nDel;
        lIns += t->lIns; nIns += t->nIns; nUnique += t->nUnique;
    }
};

struct Index {  // the slice of `class Genome` the path reads (Genome.h:26-56)
    const star_index_view_t* v;
    const char* G;
    uint nGenome, nSA;
    unsigned GstrandBit;
    uint GstrandMask, SAiMarkAbsentMaskC, SAiMarkNmaskC, SAiMarkNmask;
    unsigned saBits, saiBits;
    std::vector<uint> chrBin;
    const uint* genomeSAindexStart;
    // 2nd stage of --outFilterType BySJout: the novel junctions that passed the filters (Parameters.h:354 sjNovelN/Start/End)
    bool sjNovelOn = false;
    std::vector<uint> sjNovelStart, sjNovelEnd;

    // PackedArray::operator[] PackedArray.h:24-32
    static inline uint packed(const uint8_t* a, unsigned bits, uint ii) {
        uint b = ii * bits, B = b / 8, S = b % 8;
        uint a1;
        memcpy(&a1, a + B, 8);
        unsigned comp = 64 - bits;
        return ((a1 >> S) << comp) >> comp;
    }
    inline uint SA(uint i) const { return packed(v->SA, saBits, i); }
    inline uint SAi(uint i) const { return packed(v->SAi, saiBits, i); }

    void init(const star_index_view_t* view) {
        v = view;
        G = (const char*)view->G;
        nGenome = view->nGenome; nSA = view->nSA; GstrandBit = view->GstrandBit;
        saBits = GstrandBit + 1; saiBits = GstrandBit + 3;
        GstrandMask = ~(1ULL << GstrandBit);                       // Genome_genomeLoad.cpp:153
        SAiMarkNmaskC = 1ULL << (GstrandBit + 1); SAiMarkNmask = ~SAiMarkNmaskC;   // :157-163
        SAiMarkAbsentMaskC = 1ULL << (GstrandBit + 2);
        genomeSAindexStart = (const uint*)view->genomeSAindexStart;
        // Genome::chrBinFill Genome.cpp:209-216
        uint genomeChrBinNbases = 1ULL << view->gChrBinNbits;
        uint chrBinN = view->chrStart[view->nChrReal] / genomeChrBinNbases + 1;
        chrBin.resize(chrBinN);
        for (uint ii = 0, ichr = 1; ii < chrBinN; ++ii) {
            if (ii * genomeChrBinNbases >= view->chrStart[ichr]) ichr++;
            chrBin[ii] = ichr - 1;
        }
    }
};

struct Counters {
    uint64_t searches = 0, saiWords = 0, compareCalls = 0, basesExamined = 0, saEnum = 0, nodes = 0, leaves = 0;
};

// ---------------------------------------------------------------------------------------------
// binarySearch2.cpp:3-43
int binarySearch2(uint x, uint y, const uint* X, const uint* Y, int N) {
    if (N == 0 || x > X[N - 1] || x < X[0]) return -1;
    int i1 = 0, i2 = N - 1, i3 = N / 2;
    while (i2 > i1 + 1) {
        i3 = (i1 + i2) / 2;
        if (X[i3] > x) i2 = i3; else i1 = i3;
    }
    if (x == X[i1]) i3 = i1;
    else if (x == X[i2]) i3 = i2;
    else return -1;
    for (int jj = i3; jj >= 0; jj--) {
        if (x != X[jj]) break;
        else if (y == Y[jj]) return jj;
    }
    for (int jj = i3; jj < N; jj++) {
        if (x != X[jj]) return -1;
        else if (y == Y[jj]) return jj;
    }
    return -2;
}

// blocksOverlap.cpp:3-40
uint blocksOverlap(const Tr& t1, const Tr& t2) {
    uint i1 = 0, i2 = 0, nOverlap = 0;
    while (i1 < t1.nExons && i2 < t2.nExons) {
        uint rs1 = t1.exons[i1][EX_R], rs2 = t2.exons[i2][EX_R];
        uint re1 = rs1 + t1.exons[i1][EX_L], re2 = rs2 + t2.exons[i2][EX_L];
        uint gs1 = t1.exons[i1][EX_G], gs2 = t2.exons[i2][EX_G];
        if (rs1 >= re2) {
            i2++;
        } else if (rs2 >= re1) {
            i1++;
        } else if (gs1 - rs1 != gs2 - rs2) {
            if (re1 >= re2) i2++;
            if (re2 >= re1) i1++;
        } else {
            nOverlap += std::min(re1, re2) - std::max(rs1, rs2);
            if (re1 >= re2) i2++;
            if (re2 >= re1) i1++;
        }
    }
    return nOverlap;
}

// extendAlign.cpp:6-92
bool extendAlign(const char* R, const char* G, uint rStart, uint gStart, int dR, int dG, uint L, uint Lprev,
                 uint nMMprev, uint nMMmax, double pMMmax, bool extendToEnd, Tr* trA) {
    int iS, iG;
    int Score = 0, nMatch = 0, nMM = 0;
    trA->maxScore = 0;
    R = R + rStart;
    G = G + gStart;
    if (extendToEnd) {
        int iExt;
        for (iExt = 0; iExt < (int)L; iExt++) {
            iS = dR * iExt; iG = dG * iExt;
            if ((gStart + iG) == (uint)(-1) || G[iG] == 5) {
                trA->extendL = 0; trA->maxScore = -999999999; trA->nMatch = 0; trA->nMM = nMMmax + 1;
                return true;
            }
            if (R[iS] == MARK_FRAG_SPACER_BASE) break;
            if (R[iS] > 3 || G[iG] > 3) continue;
            if (G[iG] == R[iS]) { nMatch++; Score += scoreMatch; }
            else { nMM++; Score -= scoreMatch; }
        }
        if (iExt > 0) {
            trA->extendL = iExt; trA->maxScore = Score; trA->nMatch = nMatch; trA->nMM = nMM;
            return true;
        }
        return false;
    }
    for (int i = 0; i < (int)L; i++) {
        iS = dR * i; iG = dG * i;
        if ((gStart + iG) == (uint)(-1) || G[iG] == 5 || R[iS] == MARK_FRAG_SPACER_BASE) break;
        if (R[iS] > 3 || G[iG] > 3) continue;
        if (G[iG] == R[iS]) {
            nMatch++; Score += scoreMatch;
            if (Score > trA->maxScore) {
                if (nMM + nMMprev <= std::min(pMMmax * double(Lprev + i + 1), double(nMMmax))) {
                    trA->extendL = i + 1; trA->maxScore = Score; trA->nMatch = nMatch; trA->nMM = nMM;
                }
            }
        } else {
            if (nMM + nMMprev >= std::min(pMMmax * double(Lprev + L), double(nMMmax))) break;
            nMM++; Score -= scoreMatch;
        }
    }
    return trA->extendL > 0;
}

// stitchAlignToTranscript.cpp:9-415
intScore stitchAlignToTranscript(uint rAend, uint gAend, uint rBstart, uint gBstart, uint L, uint iFragB, uint sjAB,
                                 const star_params_t& P, const char* R, const Index& mapGen, Tr* trA,
                                 const uint outFilterMismatchNmaxTotal) {
    if (trA->nExons >= (uint)MAX_N_EXONS) return -1000010;
    const star_index_view_t& g = *mapGen.v;
    const char* G = mapGen.G;
    int Score = 0;

    if (sjAB != ((uint)-1) && trA->exons[trA->nExons - 1][EX_sjA] == sjAB
        && trA->exons[trA->nExons - 1][EX_iFrag] == iFragB && rBstart == rAend + 1 && gAend + 1 < gBstart) {  // :18-34
        if (g.sjdbMotif[sjAB] == 0 && (L <= g.sjdbShiftRight[sjAB] || trA->exons[trA->nExons - 1][EX_L] <= g.sjdbShiftLeft[sjAB]))
            return -1000006;
        trA->exons[trA->nExons][EX_L] = L;
        trA->exons[trA->nExons][EX_R] = rBstart;
        trA->exons[trA->nExons][EX_G] = gBstart;
        trA->canonSJ[trA->nExons - 1] = g.sjdbMotif[sjAB];
        trA->shiftSJ[trA->nExons - 1][0] = g.sjdbShiftLeft[sjAB];
        trA->shiftSJ[trA->nExons - 1][1] = g.sjdbShiftRight[sjAB];
        trA->sjAnnot[trA->nExons - 1] = 1;
        trA->sjStr[trA->nExons - 1] = g.sjdbStrand[sjAB];
        trA->nExons++;
        trA->nMatch += L;
        for (uint ii = rBstart; ii < rBstart + L; ii++) Score += scoreMatch;
        Score += P.sjdbScore;
    } else {
        trA->sjAnnot[trA->nExons - 1] = 0;
        trA->sjStr[trA->nExons - 1] = 0;
        if (trA->exons[trA->nExons - 1][EX_iFrag] == iFragB) {  // same fragment :40-350
            uint gBend = gBstart + L - 1;
            uint rBend = rBstart + L - 1;
            if (rBend <= rAend) return -1000001;
            if (gBend <= gAend && trA->exons[trA->nExons - 1][EX_iFrag] == iFragB) return -1000002;
            if (rBstart <= rAend) {
                gBstart += rAend - rBstart + 1;
                rBstart = rAend + 1;
                L = rBend - rBstart + 1;
            }
            for (uint ii = rBstart; ii <= rBend; ii++) Score += scoreMatch;

            int gGap = gBstart - gAend - 1;
            int rGap = rBstart - rAend - 1;
            uint nMatch = L, nMM = 0, Del = 0, Ins = 0, nIns = 0, nDel = 0;
            int jR = 0;
            int jCan = 999;
            uint gBstart1 = gBstart - rGap - 1;

            if (gGap == 0 && rGap == 0) {
            } else if (gGap > 0 && rGap > 0 && rGap == gGap) {  // :80-93
                for (int ii = 1; ii <= rGap; ii++) {
                    if (G[gAend + ii] < 4 && R[rAend + ii] < 4) {
                        if (R[rAend + ii] == G[gAend + ii]) { Score += scoreMatch; nMatch++; }
                        else { Score -= scoreMatch; nMM++; }
                    }
                }
            } else if (gGap > rGap) {  // deletion / junction :95-254
                nDel = 1;
                Del = gGap - rGap;
                if (Del > P.alignIntronMax && P.alignIntronMax > 0) return -1000003;
                int Score1 = 0;
                int jR1 = 1;
                do {
                    jR1--;
                    if (R[rAend + jR1] != G[gBstart1 + jR1] && G[gBstart1 + jR1] < 4 && R[rAend + jR1] == G[gAend + jR1]) Score1 -= scoreMatch;
                } while (Score1 + P.scoreStitchSJshift >= 0 && int(trA->exons[trA->nExons - 1][EX_L]) + jR1 > 1);

                int maxScore2 = -999999;
                Score1 = 0;
                int jPen = 0;
                do {
                    if (R[rAend + jR1] == G[gAend + jR1] && R[rAend + jR1] != G[gBstart1 + jR1]) Score1 += scoreMatch;
                    if (R[rAend + jR1] != G[gAend + jR1] && R[rAend + jR1] == G[gBstart1 + jR1]) Score1 -= scoreMatch;
                    int jCan1 = -1, jPen1 = 0, Score2 = Score1;
                    if (Del >= P.alignIntronMin) {
                        char d1 = G[gAend + jR1 + 1], d2 = G[gAend + jR1 + 2], a1 = G[gBstart1 + jR1 - 1], a2 = G[gBstart1 + jR1];
                        if (d1 == 2 && d2 == 3 && a1 == 0 && a2 == 2) { jCan1 = 1; }
                        else if (d1 == 1 && d2 == 3 && a1 == 0 && a2 == 1) { jCan1 = 2; }
                        else if (d1 == 2 && d2 == 1 && a1 == 0 && a2 == 2) { jCan1 = 3; jPen1 = P.scoreGapGCAG; }
                        else if (d1 == 1 && d2 == 3 && a1 == 2 && a2 == 1) { jCan1 = 4; jPen1 = P.scoreGapGCAG; }
                        else if (d1 == 0 && d2 == 3 && a1 == 0 && a2 == 1) { jCan1 = 5; jPen1 = P.scoreGapATAC; }
                        else if (d1 == 2 && d2 == 3 && a1 == 0 && a2 == 3) { jCan1 = 6; jPen1 = P.scoreGapATAC; }
                        else { jCan1 = 0; jPen1 = P.scoreGapNoncan; }
                        Score2 += jPen1;
                    }
                    if (maxScore2 < Score2) { maxScore2 = Score2; jR = jR1; jCan = jCan1; jPen = jPen1; }
                    jR1++;
                } while (jR1 < int(rBend) - int(rAend));

                uint jjL = 0, jjR = 0;
                while (gAend + jR >= jjL && G[gAend - jjL + jR] == G[gBstart1 - jjL + jR] && G[gAend - jjL + jR] < 4 && jjL <= MAX_SJ_REPEAT_SEARCH) jjL++;
                while (gAend + jjR + jR + 1 < mapGen.nGenome && G[gAend + jjR + jR + 1] == G[gBstart1 + jjR + jR + 1] && G[gAend + jjR + jR + 1] < 4 && jjR <= MAX_SJ_REPEAT_SEARCH) jjR++;

                if (jCan <= 0) {
                    jR -= jjL;
                    if (int(trA->exons[trA->nExons - 1][EX_L]) + jR < 1) return -1000005;
                    jjR += jjL;
                    jjL = 0;
                }
                for (int ii = std::min(1, jR + 1); ii <= std::max(rGap, jR); ii++) {
                    uint g1 = (ii <= jR) ? (gAend + ii) : (gBstart1 + ii);
                    if (G[g1] < 4 && R[rAend + ii] < 4) {
                        if (R[rAend + ii] == G[g1]) {
                            if (ii >= 1 && ii <= rGap) { Score += scoreMatch; nMatch++; }
                        } else {
                            Score -= scoreMatch; nMM++;
                            if (ii < 1 || ii > rGap) { Score -= scoreMatch; nMatch--; }
                        }
                    }
                }
                if (g.sjdbN > 0) {
                    uint jS = gAend + jR + 1, jE = gBstart1 + jR;
                    int sjdbInd = binarySearch2(jS, jE, (const uint*)g.sjdbStart, (const uint*)g.sjdbEnd, (int)g.sjdbN);
                    if (sjdbInd < 0) {
                        if (Del >= P.alignIntronMin) {
                            Score += P.scoreGap + jPen;
                        } else {
                            Score += Del * P.scoreDelBase + P.scoreDelOpen;
                            jCan = -1;
                            trA->sjAnnot[trA->nExons - 1] = 0;
                        }
                    } else {
                        jCan = g.sjdbMotif[sjdbInd];
                        if (g.sjdbMotif[sjdbInd] == 0) {
                            if (L <= g.sjdbShiftLeft[sjdbInd] || trA->exons[trA->nExons - 1][EX_L] <= g.sjdbShiftLeft[sjdbInd]) return -1000006;
                            jR += (int)g.sjdbShiftLeft[sjdbInd];
                            if (rAend + jR >= rBend) return -1000006;
                            jjL = g.sjdbShiftLeft[sjdbInd];
                            jjR = g.sjdbShiftRight[sjdbInd];
                        }
                        trA->sjAnnot[trA->nExons - 1] = 1;
                        trA->sjStr[trA->nExons - 1] = g.sjdbStrand[sjdbInd];
                        Score += P.sjdbScore;
                    }
                } else {
                    if (Del >= P.alignIntronMin) {
                        Score += P.scoreGap + jPen;
                    } else {
                        Score += Del * P.scoreDelBase + P.scoreDelOpen;
                        jCan = -1;
                        trA->sjAnnot[trA->nExons - 1] = 0;
                    }
                }
                trA->shiftSJ[trA->nExons - 1][0] = jjL;
                trA->shiftSJ[trA->nExons - 1][1] = jjR;
                trA->canonSJ[trA->nExons - 1] = jCan;
                if (trA->sjAnnot[trA->nExons - 1] == 0) {
                    if (jCan > 0) trA->sjStr[trA->nExons - 1] = 2 - jCan % 2;
                    else trA->sjStr[trA->nExons - 1] = 0;
                }
            } else if (rGap > gGap) {  // insertion :255-305
                Ins = rGap - gGap;
                nIns = 1;
                if (gGap == 0) {
                    jR = 0;
                } else if (gGap < 0) {
                    jR = 0;
                    for (int ii = 0; ii < -gGap; ii++) Score -= scoreMatch;
                } else {
                    int Score1 = 0, maxScore1 = 0;
                    for (int jR1 = 1; jR1 <= gGap; jR1++) {
                        if (G[gAend + jR1] < 4) {
                            Score1 += (R[rAend + jR1] == G[gAend + jR1]) ? scoreMatch : -scoreMatch;
                            Score1 += (R[rAend + Ins + jR1] == G[gAend + jR1]) ? -scoreMatch : +scoreMatch;
                        }
                        if (Score1 > maxScore1 || (Score1 == maxScore1 && P.alignInsertionFlushRight)) { maxScore1 = Score1; jR = jR1; }
                    }
                    for (int ii = 1; ii <= gGap; ii++) {
                        uint r1 = rAend + ii + (ii <= jR ? 0 : Ins);
                        if (G[gAend + ii] < 4 && R[r1] < 4) {
                            if (R[r1] == G[gAend + ii]) { Score += scoreMatch; nMatch++; }
                            else { Score -= scoreMatch; nMM++; }
                        }
                    }
                }
                if (P.alignInsertionFlushRight) {
                    for (; jR < (int)rBend - (int)rAend - (int)Ins; jR++) {
                        if (R[rAend + jR + 1] != G[gAend + jR + 1] || G[gAend + jR + 1] == 4) break;
                    }
                    if (jR == (int)rBend - (int)rAend - (int)Ins) return -1000009;
                }
                Score += Ins * P.scoreInsBase + P.scoreInsOpen;
                jCan = -2;
            }

            if ((trA->nMM + nMM) <= outFilterMismatchNmaxTotal
                && (jCan < 0 || (jCan < 7 && nMM <= (uint)P.alignSJstitchMismatchNmax[(jCan + 1) / 2]))) {  // :314-315
                trA->nMM += nMM;
                trA->nMatch += nMatch;
                if (Del >= P.alignIntronMin) { trA->nGap += nDel; trA->lGap += Del; }
                else { trA->nDel += nDel; trA->lDel += Del; }
                if (Del == 0 && Ins == 0) {
                    trA->exons[trA->nExons - 1][EX_L] += rBend - rAend;
                } else if (Del > 0) {
                    trA->exons[trA->nExons - 1][EX_L] += jR;
                    trA->exons[trA->nExons][EX_L] = rBend - rAend - jR;
                    trA->exons[trA->nExons][EX_R] = rAend + jR + 1;
                    trA->exons[trA->nExons][EX_G] = gBstart1 + jR + 1;
                    trA->nExons++;
                } else if (Ins > 0) {
                    trA->nIns += nIns;
                    trA->lIns += Ins;
                    trA->exons[trA->nExons - 1][EX_L] += jR;
                    trA->exons[trA->nExons][EX_L] = rBend - rAend - jR - Ins;
                    trA->exons[trA->nExons][EX_R] = rAend + jR + Ins + 1;
                    trA->exons[trA->nExons][EX_G] = gAend + 1 + jR;
                    trA->canonSJ[trA->nExons - 1] = -2;
                    trA->sjAnnot[trA->nExons - 1] = 0;
                    trA->nExons++;
                }
            } else {
                return -1000007;
            }
        } else if (gBstart + trA->exons[0][EX_R] + P.alignEndsProtrudeNbasesMax >= trA->exons[0][EX_G] || trA->exons[0][EX_G] < trA->exons[0][EX_R]) {  // mates :352-405
            if (P.alignMatesGapMax > 0 && gBstart > trA->exons[trA->nExons - 1][EX_G] + trA->exons[trA->nExons - 1][EX_L] + P.alignMatesGapMax) return -1000004;
            for (uint ii = rBstart; ii < rBstart + L; ii++) Score += scoreMatch;
            Tr trExtend;
            trExtend.reset();
            if (extendAlign(R, G, rAend + 1, gAend + 1, 1, 1, DEF_readSeqLengthMax, trA->nMatch, trA->nMM, outFilterMismatchNmaxTotal,
                            P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[trA->exons[trA->nExons - 1][EX_iFrag]][1], &trExtend)) {
                trA->add(&trExtend);
                Score += trExtend.maxScore;
                trA->exons[trA->nExons - 1][EX_L] += trExtend.extendL;
            }
            trA->exons[trA->nExons][EX_R] = rBstart;
            trA->exons[trA->nExons][EX_G] = gBstart;
            trA->exons[trA->nExons][EX_L] = L;
            trA->nMatch += L;
            trExtend.reset();
            uint extlen = P.alignEndsTypeExt[iFragB][1] ? DEF_readSeqLengthMax : gBstart - trA->exons[0][EX_G] + trA->exons[0][EX_R];
            if (extendAlign(R, G, rBstart - 1, gBstart - 1, -1, -1, extlen, trA->nMatch, trA->nMM, outFilterMismatchNmaxTotal,
                            P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[iFragB][1], &trExtend)) {
                trA->add(&trExtend);
                Score += trExtend.maxScore;
                trA->exons[trA->nExons][EX_R] -= trExtend.extendL;
                trA->exons[trA->nExons][EX_G] -= trExtend.extendL;
                trA->exons[trA->nExons][EX_L] += trExtend.extendL;
            }
            trA->canonSJ[trA->nExons - 1] = -3;
            trA->sjAnnot[trA->nExons - 1] = 0;
            trA->nExons++;
        } else {
            return -1000008;
        }
    }
    trA->exons[trA->nExons - 1][EX_iFrag] = iFragB;
    trA->exons[trA->nExons - 1][EX_sjA] = sjAB;
    return Score;
}

// ---------------------------------------------------------------------------------------------
// per-read aligner state = the members of `class ReadAlign` the path uses (ReadAlign.h:21-200)
struct ReadAlign {
    const star_params_t& P;
    const Index& mapGen;
    Counters cnt;

    char* Read1[3];
    std::vector<char> readBuf;
    uint Lread, readLength[2], readNmates;
    uint outFilterMismatchNmaxTotal;
    intScore maxScoreMate[2];

    uint splitR[3][16];
    uint Nsplit;
    std::vector<unsigned short> winBinStore;
    unsigned short* winBin[2];
    std::vector<uint> PCs, WCs, WAs;
    std::vector<uint> nWA, nWAP, WALrec, WlastAnchor;
    std::vector<char> WAincl;
    uint nW, nWall, nP, nA, nUM[2], mapMarker;
    uint multNmin, multNminL, multLmax, multLmaxN, multNmax, multNmaxL, uniqLmax, uniqLmaxInd, storedLmin;
    bool revertStrand;
    int fatal;             // STAR_EXIT_* raised inside the read
    std::string fatalMsg;

    std::vector<Tr> trArray;
    std::vector<Tr*> trArrayPointer;
    std::vector<Tr**> trAll;
    std::vector<uint> nWinTr;
    Tr trInitStore, *trInit, *trBest, trA;
    std::vector<Tr*> trMult;
    uint nTr;
    intScore maxScore;
    int unmapType;

    inline uint* PC(uint i) { return &PCs[i * PC_SIZE]; }
    inline uint* WC(uint i) { return &WCs[i * WC_SIZE]; }
    inline uint* WA(uint iW, uint iA) { return &WAs[(iW * P.seedPerWindowNmax + iA) * WA_SIZE]; }

    ReadAlign(const star_params_t& Pin, const Index& gin) : P(Pin), mapGen(gin) {  // ReadAlign.cpp:6-110
        readBuf.assign(3 * (DEF_readSeqLengthMax + 8), 0);
        for (int i = 0; i < 3; i++) Read1[i] = &readBuf[i * (DEF_readSeqLengthMax + 8)];
        winBinStore.assign(2 * P.winBinN + 2, 65535);
        winBin[0] = &winBinStore[0];
        winBin[1] = &winBinStore[P.winBinN + 1];
        PCs.resize((P.seedPerReadNmax + 2) * PC_SIZE);
        WCs.resize(P.alignWindowsPerReadNmax * WC_SIZE);
        nWA.resize(P.alignWindowsPerReadNmax); nWAP.resize(P.alignWindowsPerReadNmax);
        WALrec.resize(P.alignWindowsPerReadNmax); WlastAnchor.resize(P.alignWindowsPerReadNmax);
        WAs.resize(P.alignWindowsPerReadNmax * P.seedPerWindowNmax * WA_SIZE);
        WAincl.resize(P.seedPerWindowNmax + 1);
        trAll.resize(P.alignWindowsPerReadNmax + 1);
        nWinTr.resize(P.alignWindowsPerReadNmax);
        trArray.resize(P.alignTranscriptsPerReadNmax);
        trArrayPointer.resize(P.alignTranscriptsPerReadNmax);
        for (uint ii = 0; ii < P.alignTranscriptsPerReadNmax; ii++) trArrayPointer[ii] = &trArray[ii];
        trInit = &trInitStore;
        memset(&trInitStore, 0, sizeof(Tr));
        memset(&trA, 0, sizeof(Tr));
        fatal = 0;
        readNmates = 1;
    }

    void raise(int code, const char* msg) { if (!fatal) { fatal = code; fatalMsg = msg; } }

    void resetN() {  // ReadAlign.cpp:112-124
        mapMarker = 0; nA = 0; nP = 0; nW = 0; nTr = 0; nUM[0] = 0; nUM[1] = 0;
        storedLmin = 0; uniqLmax = 0; uniqLmaxInd = 0; multLmax = 0; multLmaxN = 0; multNminL = 0; multNmin = 0; multNmax = 0; multNmaxL = 0;
        for (uint ii = 0; ii < readNmates; ii++) maxScoreMate[ii] = 0;
    }

    // ReadAlign_oneRead.cpp:35-78 + readLoad.cpp:50 + SequenceFuns.cpp:4-14,131-146
    bool loadRead(const char* seq0, uint l0, const char* seq1, uint l1, uint nMates) {
        readNmates = nMates;
        readLength[0] = l0; readLength[1] = nMates == 2 ? l1 : 0;
        if (nMates == 1 && l0 < 1) {   // (a mate of a pair may be empty: clipped to nothing, ClipMate_clip.cpp)
            raise(STAR_EXIT_INPUT_FILES, "EXITING because of FATAL ERROR in reads input: short read sequence line: 0\n"); return false; }
        Lread = nMates == 2 ? l0 + l1 + 1 : l0;
        if (Lread > DEF_readSeqLengthMax) { raise(STAR_EXIT_INPUT_FILES, "EXITING because of FATAL ERROR in reads input: Lread of the pair exceeds DEF_readSeqLengthMax\n"); return false; }
        auto conv = [](char c) -> char {
            switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
        };
        for (uint i = 0; i < l0; i++) Read1[0][i] = conv(seq0[i]);
        if (nMates == 2) {
            Read1[0][l0] = MARK_FRAG_SPACER_BASE;
            for (uint j = 0; j < l1; j++) {
                char c = conv(seq1[l1 - 1 - j]);
                Read1[0][l0 + 1 + j] = c < 4 ? 3 - c : c;
            }
        }
        for (uint i = 0; i < Lread; i++) {
            char c = Read1[0][i];
            Read1[1][i] = c < 4 ? 3 - c : c;
        }
        for (uint i = 0; i < Lread; i++) Read1[2][Lread - i - 1] = Read1[1][i];
        outFilterMismatchNmaxTotal = std::min((uint)P.outFilterMismatchNmax, (uint)(P.outFilterMismatchNoverReadLmax * (readLength[0] + readLength[1])));
        return true;
    }

    // SequenceFuns.cpp:411-444
    uint qualitySplit(char* r, uint L, uint maxNsplit, uint minLsplit) {
        uint iR = 0, iS = 0, iR1, LgoodMin = 0, iFrag = 0;
        while ((iR < L) & (iS < maxNsplit)) {
            while (iR < L && r[iR] > 3) {
                if (r[iR] == MARK_FRAG_SPACER_BASE) iFrag++;
                iR++;
            }
            if (iR == L) break;
            iR1 = iR;
            while (iR < L && r[iR] <= 3) iR++;
            if ((iR - iR1) > LgoodMin) LgoodMin = iR - iR1;
            if ((iR - iR1) < minLsplit) continue;
            splitR[0][iS] = iR1; splitR[1][iS] = iR - iR1; splitR[2][iS] = iFrag;
            iS++;
        }
        if (iS == 0) splitR[1][0] = LgoodMin;
        return iS;
    }

    // SuffixArrayFuns.cpp:10-104
    uint compareSeqToGenome(uint S, uint N, uint L, uint iSA, bool dirR, bool& compRes) {
        long long ii;
        uint SAstr = mapGen.SA(iSA);
        bool dirG = (SAstr >> mapGen.GstrandBit) == 0;
        SAstr &= mapGen.GstrandMask;
        const char* g = mapGen.G;
        cnt.compareCalls++;
        uint ret;
        if (dirR && dirG) {
            const char* s = Read1[0] + S + L;
            g += SAstr + L;
            for (ii = 0; (uint)ii < N - L; ii++) {
                if (s[ii] != g[ii]) {
                    compRes = s[ii] > g[ii];
                    cnt.basesExamined += ii + 1;
                    return ii + L;
                }
            }
            ret = N;
        } else if (dirR && !dirG) {
            const char* s = Read1[1] + S + L;
            g += mapGen.nGenome - 1 - SAstr - L;
            for (ii = 0; (uint)ii < N - L; ii++) {
                if (s[ii] != g[-ii]) {
                    compRes = !(s[ii] > g[-ii] || g[-ii] > 3);
                    cnt.basesExamined += ii + 1;
                    return ii + L;
                }
            }
            ret = N;
        } else if (!dirR && dirG) {
            const char* s = Read1[1] + S - L;
            g += SAstr + L;
            for (ii = 0; (uint)ii < N - L; ii++) {
                if (s[-ii] != g[ii]) {
                    compRes = s[-ii] > g[ii];
                    cnt.basesExamined += ii + 1;
                    return ii + L;
                }
            }
            ret = N;
        } else {
            const char* s = Read1[0] + S - L;
            g += mapGen.nGenome - 1 - SAstr - L;
            for (ii = 0; (uint)ii < N - L; ii++) {
                if (s[-ii] != g[-ii]) {
                    compRes = !(s[-ii] > g[-ii] || g[-ii] > 3);
                    cnt.basesExamined += ii + 1;
                    return ii + L;
                }
            }
            ret = N;
        }
        cnt.basesExamined += N - L;
        return ret;
    }

    static inline uint medianUint2(uint a, uint b) { return a / 2 + b / 2 + (a % 2 + b % 2) / 2; }  // SuffixArrayFuns.cpp:4-8

    // SuffixArrayFuns.cpp:106-131
    uint findMultRange(uint i3, uint L3, uint i1, uint L1, uint i1a, uint L1a, uint i1b, uint L1b, bool dirR, uint S) {
        bool compRes;
        if (L1 < L3) {
            L1b = L1; i1b = i1; i1a = i3;
        } else {
            if (L1a < L1) { L1b = L1a; i1b = i1a; i1a = i1; }
        }
        while ((i1b + 1 < i1a) | (i1b > i1a + 1)) {
            uint i1c = medianUint2(i1a, i1b);
            uint L1c = compareSeqToGenome(S, L3, L1b, i1c, dirR, compRes);
            if (L1c == L3) i1a = i1c;
            else { i1b = i1c; L1b = L1c; }
        }
        return i1a;
    }

    // SuffixArrayFuns.cpp:133-207
    uint maxMappableLength(uint S, uint N, uint i1, uint i2, bool dirR, uint& L, uint* indStartEnd) {
        bool compRes = false;
        uint L1, L2, i3, L3, L1a, L1b, L2a, L2b, i1a, i1b, i2a, i2b;
        L1 = compareSeqToGenome(S, N, L, i1, dirR, compRes);
        L2 = compareSeqToGenome(S, N, L, i2, dirR, compRes);
        L = std::min(L1, L2);
        L1a = L1; L1b = L1; i1a = i1; i1b = i1;
        L2a = L2; L2b = L2; i2a = i2; i2b = i2;
        i3 = i1; L3 = L1;
        while (i1 + 1 < i2) {
            i3 = medianUint2(i1, i2);
            L3 = compareSeqToGenome(S, N, L, i3, dirR, compRes);
            if (L3 == N) break;
            if (compRes) {
                if (L3 > L1) { L1b = L1a; L1a = L1; i1b = i1a; i1a = i1; }
                i1 = i3; L1 = L3;
            } else {
                if (L3 > L2) { L2b = L2a; L2a = L2; i2b = i2a; i2a = i2; }
                i2 = i3; L2 = L3;
            }
            L = std::min(L1, L2);
        }
        if (L3 < N) {
            if (L1 > L2) { i3 = i1; L3 = L1; } else { i3 = i2; L3 = L2; }
        }
        i1 = findMultRange(i3, L3, i1, L1, i1a, L1a, i1b, L1b, dirR, S);
        i2 = findMultRange(i3, L3, i2, L2, i2a, L2a, i2b, L2b, dirR, S);
        L = L3;
        indStartEnd[0] = i1; indStartEnd[1] = i2;
        return i2 - i1 + 1;
    }

    // ------------------------------------------------------------------------------------------------------------------
    // Design check for the round-2 GPU seed search (test infrastructure, env STAR_ORACLE_KARY_CHECK=1): the same answer as
    // maxMappableLength — the maximal match length L over the SA rows of the start interval and the block of rows attaining it —
    // computed by a 32-ary search, i.e. the shape a warp executes cooperatively (32 probes per step instead of one, ~5x fewer
    // dependent steps).  Emulated here lane by lane on the CPU and compared with the reference binary search for every search.
    uint karyLcp(uint S, uint N, uint L, uint iSA, bool dirR, bool& compRes) {   // compareSeqToGenome without the work counters
        const auto saved = cnt;
        uint r = compareSeqToGenome(S, N, L, iSA, dirR, compRes);
        cnt = saved;
        return r;
    }
    uint karyMaxMappableLength(uint S, uint N, uint lo, uint hi, bool dirR, uint& L, uint* indStartEnd) {
        const uint K = karyK;
        uint i1 = lo, i2 = hi;
        uint L3 = 0, i3 = lo;
        uint Lc = L;                       // bases known to match for every row of [i1,i2]
        bool have = false;
        // (1) narrowing rounds: K probes spread over the window (both ends included) until the window has at most K rows
        while (i2 - i1 + 1 > K) {
            uint pr[32], pl[32]; bool pc[32];
            karyRounds++; karyProbes += K;
            for (uint j = 0; j < K; j++) {
                pr[j] = i1 + (uint)(((unsigned __int128)(i2 - i1) * j) / (K - 1));
                pl[j] = karyLcp(S, N, Lc, pr[j], dirR, pc[j]);
            }
            int jFull = -1, jLast1 = -1;
            for (uint j = 0; j < K; j++) {
                if (pl[j] == N && jFull < 0) jFull = (int)j;
                if (pl[j] != N && pc[j]) jLast1 = (int)j;   // read > suffix: the insertion point is to the right of this probe
            }
            if (jFull >= 0) { i3 = pr[jFull]; L3 = N; have = true; break; }
            if (jLast1 < 0) { i3 = pr[0]; L3 = pl[0]; have = true; break; }                    // read sorts before the first row: that row is the best
            if (jLast1 == (int)K - 1) { i3 = pr[K - 1]; L3 = pl[K - 1]; have = true; break; }  // ... after the last row
            i1 = pr[jLast1]; i2 = pr[jLast1 + 1];
            Lc = std::min(pl[jLast1], pl[jLast1 + 1]);
        }
        uint b1, b2;
        if (!have) {
            // (2) one round over every row of the window: match length per lane, maximum, block of lanes attaining it
            const uint rows = i2 - i1 + 1;
            uint pl[32] = {0}; bool pc[32];
            karyRounds++; karyProbes += rows;
            for (uint j = 0; j < rows; j++) pl[j] = karyLcp(S, N, Lc, i1 + j, dirR, pc[j]);
            uint jm = 0;
            for (uint j = 1; j < rows; j++) if (pl[j] > pl[jm]) jm = j;
            L3 = pl[jm];
            uint j1 = jm, j2 = jm;
            while (j1 > 0 && pl[j1 - 1] >= L3) j1--;
            while (j2 + 1 < rows && pl[j2 + 1] >= L3) j2++;
            b1 = i1 + j1; b2 = i1 + j2;
            i3 = i1 + jm;
        } else {
            b1 = b2 = i3;
        }
        // (3) the block may continue beyond the rows seen so far (repeats): k-ary boundary searches, clamped to [lo, hi]
        auto boundary = [&](uint inRow, uint outRow, bool left) -> uint {   // inRow has LCP >= L3; outRow (the interval end) is tested first
            bool c2;
            karyRounds++; karyProbes++;
            uint Lout = karyLcp(S, L3, L, outRow, dirR, c2);
            if (Lout >= L3) return outRow;
            uint a = outRow, La = Lout, b = inRow;        // a: LCP < L3, b: LCP >= L3
            while ((left ? a + 1 < b : b + 1 < a)) {
                const uint span = (left ? b - a : a - b) - 1;
                const uint np = std::min(K, span);
                uint pr[32], pl[32];
                karyRounds++; karyProbes += np;
                for (uint j = 0; j < np; j++) {
                    const uint step = np == span ? 1 + j : (uint)(((unsigned __int128)(span + 1) * (j + 1)) / (np + 1));
                    pr[j] = left ? a + step : a - step;
                    bool c3;
                    pl[j] = karyLcp(S, L3, La, pr[j], dirR, c3);
                }
                int jIn = -1;
                for (uint j = 0; j < np; j++) if (pl[j] >= L3) { jIn = (int)j; break; }
                if (jIn < 0) { a = pr[np - 1]; La = pl[np - 1]; }
                else { b = pr[jIn]; if (jIn > 0) { a = pr[jIn - 1]; La = pl[jIn - 1]; } }
            }
            return b;
        };
        const bool seenLeft = !have && b1 > i1, seenRight = !have && b2 < i2;   // a row with a shorter match was seen on that side
        if (!seenLeft && b1 > lo) b1 = boundary(b1, lo, true);
        if (!seenRight && b2 < hi) b2 = boundary(b2, hi, false);
        L = L3;
        indStartEnd[0] = b1; indStartEnd[1] = b2;
        return b2 - b1 + 1;
    }
    unsigned long long karyChecked = 0, karyMismatch = 0, karyRounds = 0, karyProbes = 0;
    unsigned long long winHistN[16] = {0}, winHistNodes[16] = {0};   // analysis: windows / recursion nodes by seeds per window (15 = 15+)
    uint karyK = getenv("STAR_ORACLE_KARY_K") ? std::min(32, std::max(2, atoi(getenv("STAR_ORACLE_KARY_K")))) : 32;
    bool karyCheck = getenv("STAR_ORACLE_KARY_CHECK") != nullptr;

    // ReadAlign_storeAligns.cpp:10-160 (OPTIM_STOREaligns_SIMPLE branch :27-51)
    void storeAligns(uint iDir, uint Shift, uint Nrep, uint L, uint indStartEnd[2], uint iFrag) {
        if (Nrep > P.seedMultimapNmax) {
            if (Nrep < multNmin || multNmin == 0) { multNmin = Nrep; multNminL = L; }
            return;
        }
        nUM[Nrep == 1 ? 0 : 1] += Nrep;
        nA += Nrep;
        uint rStart = iDir == 0 ? Shift : Shift + 1 - L;
        int iP;
        for (iP = nP - 1; iP >= 0; iP--) {
            if (PC(iP)[0] <= rStart) {
                if ((PC(iP)[PC_rStart] == rStart) && PC(iP)[PC_Length] < L) continue;
                if ((PC(iP)[PC_rStart] == rStart) && PC(iP)[PC_Length] == L) return;
                break;
            }
        }
        iP = iP + 1;
        for (int ii = nP - 1; ii >= iP; ii--)
            for (int jj = 0; jj < PC_SIZE; jj++) PC(ii + 1)[jj] = PC(ii)[jj];
        nP++;
        if (nP > P.seedPerReadNmax) {
            raise(STAR_EXIT_RUNTIME, "EXITING because of FATAL error: too many pieces pere read\nSOLUTION: increase input parameter --seedPerReadNmax");
            nP--;
            return;
        }
        PC(iP)[PC_rStart] = rStart; PC(iP)[PC_Length] = L; PC(iP)[PC_Str] = 0; PC(iP)[PC_Dir] = iDir; PC(iP)[PC_Nrep] = Nrep;
        PC(iP)[PC_SAstart] = indStartEnd[0]; PC(iP)[PC_SAend] = indStartEnd[1]; PC(iP)[PC_iFrag] = iFrag;
        if (L < storedLmin) L = storedLmin;
        if (Nrep == 1) {
            if (L > uniqLmax) { uniqLmax = L; uniqLmaxInd = nP - 1; }
        } else {
            if (Nrep < multNmin || multNmin == 0) { multNmin = Nrep; multNminL = L; }
            if (L > multLmax) { multLmax = L; multLmaxN = Nrep; }
            if (Nrep > multNmax) { multNmax = Nrep; multNmaxL = L; }
        }
    }

    // ReadAlign_maxMappableLength2strands.cpp:5-115 (gSAsparseD==1)
    uint maxMappableLength2strands(uint pieceStartIn, uint pieceLengthIn, uint iDir, uint iSA1, uint iSA2, uint& maxLbest, uint iFrag) {
        uint Nrep = 0, indStartEnd[2] = {0, 0}, maxL = 0;
        maxLbest = 0;
        bool dirR = iDir == 0;
        cnt.searches++;
        const uint gSAindexNbases = mapGen.v->gSAindexNbases;
        {
            uint pieceStart = pieceStartIn;
            uint pieceLength = pieceLengthIn;
            uint Lmax = std::min((uint)gSAindexNbases, pieceLength);
            uint ind1 = 0;
            if (dirR) {
                for (uint ii = 0; ii < Lmax; ii++) { ind1 <<= 2LLU; ind1 += ((uint)Read1[0][pieceStart + ii]); }
            } else {
                for (uint ii = 0; ii < Lmax; ii++) { ind1 <<= 2LLU; ind1 += (3 - ((uint)Read1[0][pieceStart - ii])); }
            }
            uint Lind = Lmax;
            while (Lind > 0) {
                iSA1 = mapGen.SAi(mapGen.genomeSAindexStart[Lind - 1] + ind1);
                cnt.saiWords++;
                if ((iSA1 & mapGen.SAiMarkAbsentMaskC) == 0) break;
                --Lind;
                ind1 = ind1 >> 2;
            }
            bool iSA2good = true;
            if (mapGen.genomeSAindexStart[Lind - 1] + ind1 + 1 < mapGen.genomeSAindexStart[Lind]) {
                iSA2 = mapGen.SAi(mapGen.genomeSAindexStart[Lind - 1] + ind1 + 1);
                cnt.saiWords++;
                if ((iSA2 & mapGen.SAiMarkAbsentMaskC) == 0) {
                    iSA2 = (iSA2 & mapGen.SAiMarkNmask) - 1;
                } else {
                    iSA2 = mapGen.nSA - 1;
                    iSA2good = false;
                }
            } else {
                iSA2 = mapGen.nSA - 1;
                iSA2good = false;
            }
            bool iSA1noN = (iSA1 & mapGen.SAiMarkNmaskC) == 0;
            if (Lind < gSAindexNbases && iSA1noN && iSA2good) {
                indStartEnd[0] = iSA1; indStartEnd[1] = iSA2;
                Nrep = indStartEnd[1] - indStartEnd[0] + 1;
                maxL = Lind;
            } else if (iSA1 == iSA2 && iSA1noN && iSA2good) {
                indStartEnd[0] = indStartEnd[1] = iSA1;
                Nrep = 1;
                bool comparRes;
                maxL = compareSeqToGenome(pieceStart, pieceLength, Lind, iSA1, dirR, comparRes);
            } else {
                if (iSA2good && iSA1noN) maxL = Lind; else maxL = 0;
                uint kL = maxL, kInd[2] = {0, 0}, kN = 0;
                if (karyCheck) kN = karyMaxMappableLength(pieceStart, pieceLength, iSA1 & mapGen.SAiMarkNmask, iSA2, dirR, kL, kInd);
                Nrep = maxMappableLength(pieceStart, pieceLength, iSA1 & mapGen.SAiMarkNmask, iSA2, dirR, maxL, indStartEnd);
                if (karyCheck) {
                    karyChecked++;
                    if (kN != Nrep || kL != maxL || kInd[0] != indStartEnd[0] || kInd[1] != indStartEnd[1]) {
                        if (karyMismatch++ < 5)
                            fprintf(stderr, "KARY MISMATCH: ref L=%llu [%llu,%llu] kary L=%llu [%llu,%llu] (S=%llu N=%llu dir=%d)\n", maxL, indStartEnd[0], indStartEnd[1], kL, kInd[0], kInd[1], pieceStart, pieceLength, (int)dirR);
                    }
                }
            }
            if (maxL > maxLbest) maxLbest = maxL;
        }
        storeAligns(iDir, pieceStartIn, Nrep, maxL, indStartEnd, iFrag);
        return Nrep;
    }

    // sjAlignSplit.cpp:3-15
    bool sjAlignSplit(uint a1, uint aLength, uint& a1D, uint& aLengthD, uint& a1A, uint& aLengthA, uint& isj) {
        const star_index_view_t& g = *mapGen.v;
        uint sj1 = (a1 - g.sjGstart) % g.sjdbLength;
        if (sj1 < g.sjdbOverhang && sj1 + aLength > g.sjdbOverhang) {
            isj = (a1 - g.sjGstart) / g.sjdbLength;
            aLengthD = g.sjdbOverhang - sj1;
            aLengthA = aLength - aLengthD;
            a1D = g.sjDstart[isj] + sj1;
            a1A = g.sjAstart[isj];
            return true;
        }
        return false;
    }

    // ReadAlign_createExtendWindowsWithAlign.cpp:7-84
    int createExtendWindowsWithAlign(uint a1, uint aStr) {
        uint aBin = (a1 >> P.winBinNbits);
        uint iBinLeft = aBin, iBinRight = aBin;
        unsigned short* wB = winBin[aStr];
        uint iBin = -1, iWin = -1, iWinRight = -1;
        if (wB[aBin] == uintWinBinMax) {
            bool flagMergeLeft = false;
            if (aBin > 0) {
                for (iBin = aBin - 1; iBin >= (aBin > P.winAnchorDistNbins ? aBin - P.winAnchorDistNbins : 0); --iBin) {
                    if (wB[iBin] < uintWinBinMax) { flagMergeLeft = true; break; }
                    if (iBin == 0) break;
                }
                flagMergeLeft = flagMergeLeft && (mapGen.chrBin[iBin >> P.winBinChrNbits] == mapGen.chrBin[aBin >> P.winBinChrNbits]);
                if (flagMergeLeft) {
                    iWin = wB[iBin];
                    iBinLeft = WC(iWin)[WC_gStart];
                    for (uint ii = iBin + 1; ii <= aBin; ii++) wB[ii] = iWin;
                }
            }
            bool flagMergeRight = false;
            if (aBin + 1 < P.winBinN) {
                for (iBin = aBin + 1; iBin < std::min(aBin + P.winAnchorDistNbins + 1, (uint)P.winBinN); ++iBin) {
                    if (wB[iBin] < uintWinBinMax) { flagMergeRight = true; break; }
                }
                flagMergeRight = flagMergeRight && (mapGen.chrBin[iBin >> P.winBinChrNbits] == mapGen.chrBin[aBin >> P.winBinChrNbits]);
                if (flagMergeRight) {
                    while (wB[iBin] == wB[iBin + 1]) ++iBin;
                    iBinRight = iBin;
                    iWinRight = wB[iBin];
                    if (!flagMergeLeft) iWin = wB[iBin];
                    for (uint ii = aBin; ii <= iBin; ii++) wB[ii] = iWin;
                }
            }
            if (!flagMergeLeft && !flagMergeRight) {
                wB[aBin] = iWin = nW;
                WC(iWin)[WC_Chr] = mapGen.chrBin[aBin >> P.winBinChrNbits];
                WC(iWin)[WC_Str] = aStr;
                WC(iWin)[WC_gEnd] = WC(iWin)[WC_gStart] = aBin;
                ++nW;
                if (nW >= P.alignWindowsPerReadNmax) {
                    nW = P.alignWindowsPerReadNmax - 1;
                    return TOO_MANY_WINDOWS;
                }
            } else {
                WC(iWin)[WC_gStart] = iBinLeft;
                WC(iWin)[WC_gEnd] = iBinRight;
                if (flagMergeLeft && flagMergeRight) {
                    WC(iWinRight)[WC_gStart] = 1;
                    WC(iWinRight)[WC_gEnd] = 0;
                }
            }
        }
        return 0;
    }

    // ReadAlign_assignAlignToWindow.cpp:6-130
    void assignAlignToWindow(uint a1, uint aLength, uint aStr, uint aNrep, uint aFrag, uint aRstart, bool aAnchor, uint sjA) {
        uint iW = winBin[aStr][a1 >> P.winBinNbits];
        if (iW == uintWinBinMax || (!aAnchor && aLength < WALrec[iW])) return;
        {
            uint iA;
            for (iA = 0; iA < nWA[iW]; iA++) {
                if (aFrag == WA(iW, iA)[WA_iFrag] && WA(iW, iA)[WA_sjA] == sjA
                    && a1 + WA(iW, iA)[WA_rStart] == WA(iW, iA)[WA_gStart] + aRstart
                    && ((aRstart >= WA(iW, iA)[WA_rStart] && aRstart < WA(iW, iA)[WA_rStart] + WA(iW, iA)[WA_Length])
                        || (aRstart + aLength >= WA(iW, iA)[WA_rStart] && aRstart + aLength < WA(iW, iA)[WA_rStart] + WA(iW, iA)[WA_Length]))) {
                    break;
                }
            }
            if (iA < nWA[iW]) {
                if (aLength > WA(iW, iA)[WA_Length]) {
                    uint iA0;
                    for (iA0 = 0; iA0 < nWA[iW]; iA0++) {
                        if (iA0 != iA && aRstart < WA(iW, iA0)[WA_rStart]) break;
                    }
                    if (iA0 > iA) --iA0;
                    if (iA0 < iA) {
                        for (uint iA1 = iA; iA1 > iA0; iA1--)
                            for (uint ii = 0; ii < WA_SIZE; ii++) WA(iW, iA1)[ii] = WA(iW, iA1 - 1)[ii];
                    } else if (iA0 > iA) {
                        for (uint iA1 = iA; iA1 < iA0; iA1++)
                            for (uint ii = 0; ii < WA_SIZE; ii++) WA(iW, iA1)[ii] = WA(iW, iA1 + 1)[ii];
                    }
                    WA(iW, iA0)[WA_rStart] = aRstart; WA(iW, iA0)[WA_Length] = aLength; WA(iW, iA0)[WA_gStart] = a1;
                    WA(iW, iA0)[WA_Nrep] = aNrep; WA(iW, iA0)[WA_Anchor] = int(aAnchor); WA(iW, iA0)[WA_iFrag] = aFrag; WA(iW, iA0)[WA_sjA] = sjA;
                }
                return;
            }
        }
        if (nWA[iW] == P.seedPerWindowNmax) {
            WALrec[iW] = Lread + 1;
            for (uint iA = 0; iA < nWA[iW]; iA++)
                if (WA(iW, iA)[WA_Anchor] != 1) WALrec[iW] = std::min(WALrec[iW], WA(iW, iA)[WA_Length]);
            if (WALrec[iW] == Lread + 1) {
                mapMarker = STAR_MARKER_TOO_MANY_ANCHORS_PER_WINDOW;
                nW = 0;
                return;
            }
            if (!aAnchor && aLength < WALrec[iW]) return;
            uint iA1 = 0;
            for (uint iA = 0; iA < nWA[iW]; iA++) {
                if (WA(iW, iA)[WA_Anchor] == 1 || WA(iW, iA)[WA_Length] > WALrec[iW]) {
                    for (uint ii = 0; ii < WA_SIZE; ii++) WA(iW, iA1)[ii] = WA(iW, iA)[ii];
                    iA1++;
                }
            }
            nWA[iW] = iA1;
            if (!aAnchor && aLength <= WALrec[iW]) nWAP[iW] = 0;
        }
        if (aAnchor || aLength > WALrec[iW]) {
            if (nWA[iW] >= P.seedPerWindowNmax) { raise(STAR_EXIT_BUG, "BUG: iA>=P.seedPerWindowNmax in stitchPieces, exiting"); return; }
            uint iA;
            for (iA = 0; iA < nWA[iW]; iA++)
                if (aRstart < WA(iW, iA)[WA_rStart]) break;
            for (uint iA1 = nWA[iW]; iA1 > iA; iA1--)
                for (uint ii = 0; ii < WA_SIZE; ii++) WA(iW, iA1)[ii] = WA(iW, iA1 - 1)[ii];
            WA(iW, iA)[WA_rStart] = aRstart; WA(iW, iA)[WA_Length] = aLength; WA(iW, iA)[WA_gStart] = a1;
            WA(iW, iA)[WA_Nrep] = aNrep; WA(iW, iA)[WA_Anchor] = int(aAnchor); WA(iW, iA)[WA_iFrag] = aFrag; WA(iW, iA)[WA_sjA] = sjA;
            nWA[iW]++;
            nWAP[iW]++;
            if (aAnchor && WlastAnchor[iW] < iA) WlastAnchor[iW] = iA;
        }
    }

    // stitchWindowAligns.cpp:8-353 (recursive; trA by value exactly like the reference)
    void stitchWindowAligns(uint iA, uint nA_, int Score, char* WAincl_, uint tR2, uint tG2, Tr trA_, uint iW, const char* R, Tr** wTr, uint* nWinTr_) {
        cnt.nodes++;
        if (iA >= nA_ && tR2 == 0) return;
        if (iA >= nA_) {
            cnt.leaves++;
            Tr& trA = trA_;
            Tr trAstep1;
            int vOrder[2];
            if (trA.roStr == 0) { vOrder[0] = 0; vOrder[1] = 1; } else { vOrder[0] = 1; vOrder[1] = 0; }
            for (int iOrd = 0; iOrd < 2; iOrd++) {
                switch (vOrder[iOrd]) {
                    case 0:
                        if (trA.rStart > 0) {
                            trAstep1.reset();
                            uint imate = trA.exons[0][EX_iFrag];
                            if (extendAlign(R, mapGen.G, trA.rStart - 1, trA.gStart - 1, -1, -1, trA.rStart, tR2 - trA.rStart + 1, trA.nMM,
                                            outFilterMismatchNmaxTotal, P.outFilterMismatchNoverLmax,
                                            P.alignEndsTypeExt[imate][(int)(trA.Str != imate)], &trAstep1)) {
                                trA.add(&trAstep1);
                                Score += trAstep1.maxScore;
                                trA.exons[0][EX_R] = trA.rStart = trA.rStart - trAstep1.extendL;
                                trA.exons[0][EX_G] = trA.gStart = trA.gStart - trAstep1.extendL;
                                trA.exons[0][EX_L] += trAstep1.extendL;
                            }
                        }
                        break;
                    case 1:
                        if (tR2 < Lread) {
                            trAstep1.reset();
                            uint imate = trA.exons[trA.nExons - 1][EX_iFrag];
                            if (extendAlign(R, mapGen.G, tR2 + 1, tG2 + 1, +1, +1, Lread - tR2 - 1, tR2 - trA.rStart + 1, trA.nMM,
                                            outFilterMismatchNmaxTotal, P.outFilterMismatchNoverLmax,
                                            P.alignEndsTypeExt[imate][(int)(imate == trA.Str)], &trAstep1)) {
                                trA.add(&trAstep1);
                                Score += trAstep1.maxScore;
                                tR2 += trAstep1.extendL;
                                tG2 += trAstep1.extendL;
                                trA.exons[trA.nExons - 1][EX_L] += trAstep1.extendL;
                            }
                        }
                }
            }
            const star_index_view_t& g = *mapGen.v;
            if (!P.alignSoftClipAtReferenceEnds &&
                ((trA.exons[trA.nExons - 1][EX_G] + Lread - trA.exons[trA.nExons - 1][EX_R]) > (g.chrStart[trA.Chr] + g.chrLength[trA.Chr]) ||
                 trA.exons[0][EX_G] < (g.chrStart[trA.Chr] + trA.exons[0][EX_R]))) {
                return;
            }
            trA.rLength = 0;
            for (uint isj = 0; isj < trA.nExons; isj++) trA.rLength += trA.exons[isj][EX_L];
            trA.gLength = tG2 + 1 - trA.gStart;

            for (uint isj = 0; isj < trA.nExons - 1; isj++) {  // :96-106
                if (trA.canonSJ[isj] >= 0) {
                    if (trA.sjAnnot[isj] == 1) {
                        if ((trA.exons[isj][EX_L] < P.alignSJDBoverhangMin && (isj == 0 || trA.canonSJ[isj - 1] == -3 || (trA.sjAnnot[isj - 1] == 0 && trA.canonSJ[isj - 1] >= 0)))
                            || (trA.exons[isj + 1][EX_L] < P.alignSJDBoverhangMin && (isj == trA.nExons - 2 || trA.canonSJ[isj + 1] == -3 || (trA.sjAnnot[isj + 1] == 0 && trA.canonSJ[isj + 1] >= 0))))
                            return;
                    } else {
                        if (trA.exons[isj][EX_L] < P.alignSJoverhangMin + trA.shiftSJ[isj][0]
                            || trA.exons[isj + 1][EX_L] < P.alignSJoverhangMin + trA.shiftSJ[isj][1]) return;
                    }
                }
            }
            if (trA.nExons > 1 && trA.sjAnnot[trA.nExons - 2] == 1 && trA.exons[trA.nExons - 1][EX_L] < P.alignSJDBoverhangMin) return;

            uint sjN = 0;  // :110-135
            trA.intronMotifs[0] = 0; trA.intronMotifs[1] = 0; trA.intronMotifs[2] = 0;
            trA.sjYes = false;
            for (uint iex = 0; iex < trA.nExons - 1; iex++) {
                if (trA.canonSJ[iex] >= 0) {
                    sjN++;
                    trA.intronMotifs[trA.sjStr[iex]]++;
                    trA.sjYes = true;
                }
            }
            if (trA.intronMotifs[1] > 0 && trA.intronMotifs[2] == 0) trA.sjMotifStrand = 1;
            else if (trA.intronMotifs[1] == 0 && trA.intronMotifs[2] > 0) trA.sjMotifStrand = 2;
            else trA.sjMotifStrand = 0;
            if (trA.intronMotifs[1] > 0 && trA.intronMotifs[2] > 0 && P.outFilterIntronStrandsRemoveInconsistent) return;
            if (sjN > 0 && trA.sjMotifStrand == 0 && P.outSAMstrandFieldType == 1) return;

            if (P.outFilterIntronMotifs == 1) {  // :137-152
                for (uint iex = 0; iex < trA.nExons - 1; iex++) if (trA.canonSJ[iex] == 0) return;
            } else if (P.outFilterIntronMotifs == 2) {
                for (uint iex = 0; iex < trA.nExons - 1; iex++) if (trA.canonSJ[iex] == 0 && trA.sjAnnot[iex] == 0) return;
            }
            {  // :154-167
                uint nsj = 0, exl = 0;
                for (uint iex = 0; iex < trA.nExons; iex++) {
                    exl += trA.exons[iex][EX_L];
                    if (iex == trA.nExons - 1 || trA.canonSJ[iex] == -3) {
                        if (nsj > 0 && (exl < P.alignSplicedMateMapLmin || exl < (uint)(P.alignSplicedMateMapLminOverLmate * readLength[trA.exons[iex][EX_iFrag]]))) return;
                        exl = 0; nsj = 0;
                    } else if (trA.canonSJ[iex] >= 0) {
                        nsj++;
                    }
                }
            }
            if (mapGen.sjNovelOn) {  // :169-177 outFilterBySJoutStage==2: unannotated junctions have to be in the filtered set
                for (uint iex = 0; iex < trA.nExons - 1; iex++) {
                    if (trA.canonSJ[iex] >= 0 && trA.sjAnnot[iex] == 0) {
                        uint jS = trA.exons[iex][EX_G] + trA.exons[iex][EX_L];
                        uint jE = trA.exons[iex + 1][EX_G] - 1;
                        if (binarySearch2(jS, jE, mapGen.sjNovelStart.data(), mapGen.sjNovelEnd.data(), (int)mapGen.sjNovelStart.size()) < 0) return;
                    }
                }
            }
            if (trA.exons[0][EX_iFrag] != trA.exons[trA.nExons - 1][EX_iFrag]) {  // :179-219
                if (trA.exons[trA.nExons - 1][EX_G] + trA.exons[trA.nExons - 1][EX_L] <= trA.exons[0][EX_G]) return;
                uint iexM2 = trA.nExons;
                for (uint iex = 0; iex < trA.nExons - 1; iex++) {
                    if (trA.canonSJ[iex] == -3) { iexM2 = iex + 1; break; }
                }
                if (trA.exons[iexM2 - 1][EX_G] + trA.exons[iexM2 - 1][EX_L] > trA.exons[iexM2][EX_G]) {
                    if (trA.exons[0][EX_G] > trA.exons[iexM2][EX_G] + trA.exons[0][EX_R] + P.alignEndsProtrudeNbasesMax) return;
                    if (trA.exons[iexM2 - 1][EX_G] + trA.exons[iexM2 - 1][EX_L] > trA.exons[trA.nExons - 1][EX_G] + Lread - trA.exons[trA.nExons - 1][EX_R] + P.alignEndsProtrudeNbasesMax) return;
                    uint iex1 = 1, iex2 = iexM2 + 1;
                    for (; iex1 < iexM2; iex1++) {
                        if (trA.exons[iex1][EX_G] >= trA.exons[iex2 - 1][EX_G] + trA.exons[iex2 - 1][EX_L]) break;
                    }
                    while (iex1 < iexM2 && iex2 < trA.nExons) {
                        if (trA.canonSJ[iex1 - 1] < 0) { iex1++; continue; }
                        if (trA.canonSJ[iex2 - 1] < 0) { iex2++; continue; }
                        if ((trA.exons[iex1][EX_G] != trA.exons[iex2][EX_G]) || ((trA.exons[iex1 - 1][EX_G] + trA.exons[iex1 - 1][EX_L]) != (trA.exons[iex2 - 1][EX_G] + trA.exons[iex2 - 1][EX_L]))) return;
                        iex1++; iex2++;
                    }
                }
            }
            if (P.scoreGenomicLengthLog2scale != 0) {  // :221-225
                Score += int(std::ceil(std::log2((double)(trA.exons[trA.nExons - 1][EX_G] + trA.exons[trA.nExons - 1][EX_L] - trA.exons[0][EX_G])) * P.scoreGenomicLengthLog2scale - 0.5));
                Score = std::max(0, Score);
            }
            trA.roStart = (trA.roStr == 0) ? trA.rStart : Lread - trA.rStart - trA.rLength;
            trA.maxScore = Score;
            if (trA.exons[0][EX_iFrag] == trA.exons[trA.nExons - 1][EX_iFrag]) {
                trA.iFrag = trA.exons[0][EX_iFrag];
                maxScoreMate[trA.iFrag] = std::max(maxScoreMate[trA.iFrag], Score);
            } else {
                trA.iFrag = -1;
            }
            trA.maxScore = Score;
            if (Score + P.outFilterMultimapScoreRange >= wTr[0]->maxScore
                || (trA.iFrag >= 0 && Score + P.outFilterMultimapScoreRange >= maxScoreMate[trA.iFrag])) {  // :245-247 (pCh.segmentMin==0)
                uint iTr = 0;
                trA.mappedLength = 0;
                for (uint iex = 0; iex < trA.nExons; iex++) trA.mappedLength += trA.exons[iex][EX_L];
                while (iTr < *nWinTr_) {
                    uint nOverlap = blocksOverlap(trA, *wTr[iTr]);
                    uint uNew = trA.mappedLength - nOverlap;
                    uint uOld = wTr[iTr]->mappedLength - nOverlap;
                    if (uNew == 0 && Score < wTr[iTr]->maxScore) {
                        break;
                    } else if (uOld == 0) {
                        Tr* pTr = wTr[iTr];
                        for (uint ii = iTr + 1; ii < *nWinTr_; ii++) wTr[ii - 1] = wTr[ii];
                        (*nWinTr_)--;
                        wTr[*nWinTr_] = pTr;
                    } else if (uOld > 0 && (uNew > 0 || Score >= wTr[iTr]->maxScore)) {
                        iTr++;
                    }
                }
                if (iTr == *nWinTr_) {
                    for (iTr = 0; iTr < *nWinTr_; iTr++) {
                        if (Score > wTr[iTr]->maxScore || (Score == wTr[iTr]->maxScore && trA.gLength < wTr[iTr]->gLength)) break;
                    }
                    Tr* pTr = wTr[*nWinTr_];
                    for (int ii = *nWinTr_; ii > int(iTr); ii--) wTr[ii] = wTr[ii - 1];
                    wTr[iTr] = pTr;
                    *(wTr[iTr]) = trA;
                    if (*nWinTr_ < P.alignTranscriptsPerWindowNmax) (*nWinTr_)++;
                }
            }
            return;
        }
        // ------------------------------------------------------------------ :308-352
        int dScore = 0;
        Tr trAi = trA_;
        uint* wa = WA(iW, iA);
        if (trA_.nExons > 0) {
            dScore = stitchAlignToTranscript(tR2, tG2, wa[WA_rStart], wa[WA_gStart], wa[WA_Length], wa[WA_iFrag], wa[WA_sjA], P, R, mapGen, &trAi, outFilterMismatchNmaxTotal);
        } else {
            trAi.exons[0][EX_R] = trAi.rStart = wa[WA_rStart];
            trAi.exons[0][EX_G] = trAi.gStart = wa[WA_gStart];
            trAi.exons[0][EX_L] = wa[WA_Length];
            trAi.exons[0][EX_iFrag] = wa[WA_iFrag];
            trAi.exons[0][EX_sjA] = wa[WA_sjA];
            trAi.nExons = 1;
            for (uint ii = 0; ii < wa[WA_Length]; ii++) dScore += scoreMatch;
            trAi.nMatch = wa[WA_Length];
            for (uint ii = 0; ii < nA_; ii++) WAincl_[ii] = false;
        }
        if (dScore > -1000000) {
            WAincl_[iA] = true;
            if (wa[WA_Nrep] == 1) trAi.nUnique++;
            if (wa[WA_Anchor] > 0) trAi.nAnchor++;
            stitchWindowAligns(iA + 1, nA_, Score + dScore, WAincl_, wa[WA_rStart] + wa[WA_Length] - 1, wa[WA_gStart] + wa[WA_Length] - 1, trAi, iW, R, wTr, nWinTr_);
        }
        if (wa[WA_Anchor] != 2 || trA_.nAnchor > 0) {
            WAincl_[iA] = false;
            stitchWindowAligns(iA + 1, nA_, Score, WAincl_, tR2, tG2, trA_, iW, R, wTr, nWinTr_);
        }
    }

    // ReadAlign_stitchPieces.cpp:12-350
    void stitchPieces() {
        for (uint i = 0; i < P.winBinN; i++) { winBin[0][i] = 65535; winBin[1][i] = 65535; }
        nW = 0;
        for (uint iP = 0; iP < nP; iP++) {
            if (PC(iP)[PC_Nrep] <= P.winAnchorMultimapNmax) {
                uint aDir = PC(iP)[PC_Dir];
                uint aLength = PC(iP)[PC_Length];
                for (uint iSA = PC(iP)[PC_SAstart]; iSA <= PC(iP)[PC_SAend]; iSA++) {
                    cnt.saEnum++;
                    uint a1 = mapGen.SA(iSA);
                    uint aStr = a1 >> mapGen.GstrandBit;
                    a1 &= mapGen.GstrandMask;
                    if (aDir == 1 && aStr == 0) {
                        aStr = 1;
                    } else if (aDir == 0 && aStr == 1) {
                        a1 = mapGen.nGenome - (aLength + a1);
                    } else if (aDir == 1 && aStr == 1) {
                        aStr = 0;
                        a1 = mapGen.nGenome - (aLength + a1);
                    }
                    if (revertStrand) aStr = 1 - aStr;
                    if (a1 >= mapGen.v->sjGstart) {
                        uint a1D, aLengthD, a1A, aLengthA, sj1;
                        if (sjAlignSplit(a1, aLength, a1D, aLengthD, a1A, aLengthA, sj1)) {
                            int addStatus = createExtendWindowsWithAlign(a1D, aStr);
                            if (addStatus == TOO_MANY_WINDOWS) break;
                            addStatus = createExtendWindowsWithAlign(a1A, aStr);
                            if (addStatus == TOO_MANY_WINDOWS) break;
                        }
                    } else {
                        int addStatus = createExtendWindowsWithAlign(a1, aStr);
                        if (addStatus == TOO_MANY_WINDOWS) break;
                    }
                }
            }
        }
        for (uint iWin = 0; iWin < nW; iWin++) {  // :96-118
            if (WC(iWin)[WC_gStart] <= WC(iWin)[WC_gEnd]) {
                uint wb = WC(iWin)[WC_gStart];
                for (uint ii = 0; ii < P.winFlankNbins && wb > 0 && mapGen.chrBin[(wb - 1) >> P.winBinChrNbits] == WC(iWin)[WC_Chr]; ii++) {
                    wb--;
                    winBin[WC(iWin)[WC_Str]][wb] = (unsigned short)iWin;
                }
                WC(iWin)[WC_gStart] = wb;
                wb = WC(iWin)[WC_gEnd];
                for (uint ii = 0; ii < P.winFlankNbins && wb + 1 < P.winBinN && mapGen.chrBin[(wb + 1) >> P.winBinChrNbits] == WC(iWin)[WC_Chr]; ii++) {
                    wb++;
                    winBin[WC(iWin)[WC_Str]][wb] = (unsigned short)iWin;
                }
                WC(iWin)[WC_gEnd] = wb;
            }
            nWA[iWin] = 0; WALrec[iWin] = 0; WlastAnchor[iWin] = -1;
        }
        nWall = nW;
        for (uint iP = 0; iP < nP; iP++) {  // :129-185
            uint aNrep = PC(iP)[PC_Nrep], aFrag = PC(iP)[PC_iFrag], aLength = PC(iP)[PC_Length], aDir = PC(iP)[PC_Dir];
            bool aAnchor = (aNrep <= P.winAnchorMultimapNmax);
            for (uint ii = 0; ii < nW; ii++) nWAP[ii] = 0;
            for (uint iSA = PC(iP)[PC_SAstart]; iSA <= PC(iP)[PC_SAend]; iSA++) {
                if (mapMarker != STAR_MARKER_TOO_MANY_ANCHORS_PER_WINDOW) cnt.saEnum++;   // after that marker the reference's enumeration is dead work
                uint a1 = mapGen.SA(iSA);
                uint aStr = a1 >> mapGen.GstrandBit;
                a1 &= mapGen.GstrandMask;
                uint aRstart = PC(iP)[PC_rStart];
                if (aDir == 1 && aStr == 0) {
                    aStr = 1;
                    aRstart = Lread - (aLength + aRstart);
                } else if (aDir == 0 && aStr == 1) {
                    aRstart = Lread - (aLength + aRstart);
                    a1 = mapGen.nGenome - (aLength + a1);
                } else if (aDir == 1 && aStr == 1) {
                    aStr = 0;
                    a1 = mapGen.nGenome - (aLength + a1);
                }
                if (revertStrand) aStr = 1 - aStr;
                if (a1 >= mapGen.v->sjGstart) {
                    uint a1D, aLengthD, a1A, aLengthA, isj1;
                    if (sjAlignSplit(a1, aLength, a1D, aLengthD, a1A, aLengthA, isj1)) {
                        assignAlignToWindow(a1D, aLengthD, aStr, aNrep, aFrag, aRstart, aAnchor, isj1);
                        assignAlignToWindow(a1A, aLengthA, aStr, aNrep, aFrag, aRstart + aLengthD, aAnchor, isj1);
                    } else {
                        continue;
                    }
                } else {
                    assignAlignToWindow(a1, aLength, aStr, aNrep, aFrag, aRstart, aAnchor, -1);
                }
            }
        }
        // :262-350
        trBest = trInit;
        uint iW1 = 0;
        uint trNtotal = 0;
        for (uint iW = 0; iW < nW; iW++) {
            if (nWA[iW] == 0) continue;
            if (WlastAnchor[iW] < nWA[iW]) WA(iW, WlastAnchor[iW])[WA_Anchor] = 2;
            for (uint ii = 0; ii < nWA[iW]; ii++) WAincl[ii] = false;
            trA = *trInit;
            trA.Chr = WC(iW)[WC_Chr];
            trA.Str = WC(iW)[WC_Str];
            trA.roStr = revertStrand ? 1 - trA.Str : trA.Str;
            trA.maxScore = 0;
            trAll[iW1] = &trArrayPointer[trNtotal];
            if (trNtotal + P.alignTranscriptsPerWindowNmax >= P.alignTranscriptsPerReadNmax) break;  // logs a warning in the reference
            *(trAll[iW1][0]) = trA;
            nWinTr[iW1] = 0;
            const unsigned long long nodes0 = cnt.nodes;
            stitchWindowAligns(0, nWA[iW], 0, &WAincl[0], 0, 0, trA, iW, Read1[trA.roStr == 0 ? 0 : 2], trAll[iW1], &nWinTr[iW1]);
            { const uint hb = nWA[iW] < 15 ? nWA[iW] : 15; winHistN[hb]++; winHistNodes[hb] += cnt.nodes - nodes0; }
            if (nWinTr[iW1] == 0) continue;
            if (trAll[iW1][0]->maxScore > trBest->maxScore || (trAll[iW1][0]->maxScore == trBest->maxScore && trAll[iW1][0]->gLength < trBest->gLength)) trBest = trAll[iW1][0];
            trNtotal += nWinTr[iW1];
            iW1++;
        }
        nW = iW1;
        if (trBest->maxScore == 0) {
            mapMarker = STAR_MARKER_NO_GOOD_WINDOW;
            nW = 0;
            return;
        }
    }

    // ReadAlign_mapOneRead.cpp:6-118
    void mapOneRead() {
        revertStrand = false;
        if (Lread > 0) Nsplit = qualitySplit(Read1[0], Lread, P.maxNsplit, P.seedSplitMin);
        else Nsplit = 0;
        resetN();
        trInit->reset();
        trInit->Chr = 0; trInit->Str = 0; trInit->roStr = 0; trInit->cStart = 0; trInit->gLength = 0;
        trInit->nExons = 0;
        trBest = trInit;
        uint seedSearchStartLmax = std::min((uint)P.seedSearchStartLmax, (uint)(P.seedSearchStartLmaxOverLread * (Lread - 1)));
        for (uint ip = 0; ip < Nsplit; ip++) {
            uint Nstart = P.seedSearchStartLmax > 0 && seedSearchStartLmax < splitR[1][ip] ? splitR[1][ip] / seedSearchStartLmax + 1 : 1;
            uint Lstart = splitR[1][ip] / Nstart;
            bool flagDirMap = true;
            for (uint iDir = 0; iDir < 2; iDir++) {
                uint Lmapped, L;
                for (uint istart = 0; istart < Nstart; istart++) {
                    if (flagDirMap || istart > 0) {
                        Lmapped = 0;
                        while (istart * Lstart + Lmapped + P.seedMapMin < splitR[1][ip]) {
                            uint Shift = iDir == 0 ? (splitR[0][ip] + istart * Lstart + Lmapped) : (splitR[0][ip] + splitR[1][ip] - istart * Lstart - 1 - Lmapped);
                            uint seedLength = splitR[1][ip] - Lmapped - istart * Lstart;
                            maxMappableLength2strands(Shift, seedLength, iDir, 0, mapGen.nSA - 1, L, splitR[2][ip]);
                            if (iDir == 0 && istart == 0 && Lmapped == 0 && Shift + L == splitR[1][ip]) flagDirMap = false;
                            Lmapped += L;
                            if (fatal) return;
                        }
                    }
                    if (P.seedSearchLmax > 0) {   // :81-87 one more search of fixed maximum length from every start
                        uint Shift = iDir == 0 ? (splitR[0][ip] + istart * Lstart) : (splitR[0][ip] + splitR[1][ip] - istart * Lstart - 1);
                        uint seedLength = std::min((uint)P.seedSearchLmax, iDir == 0 ? (splitR[0][ip] + splitR[1][ip] - Shift) : (Shift + 1));
                        maxMappableLength2strands(Shift, seedLength, iDir, 0, mapGen.nSA - 1, L, splitR[2][ip]);
                        if (fatal) return;
                    }
                }
            }
        }
        if (Lread < P.outFilterMatchNmin) {
            mapMarker = STAR_MARKER_READ_TOO_SHORT; trBest->rLength = 0; nW = 0;
        } else if (Nsplit == 0) {
            mapMarker = STAR_MARKER_NO_GOOD_PIECES; trBest->rLength = splitR[1][0]; nW = 0;
        } else if (Nsplit > 0 && nA == 0) {
            mapMarker = STAR_MARKER_ALL_PIECES_EXCEED_seedMultimapNmax; trBest->rLength = multNminL; nW = 0;
        } else if (Nsplit > 0 && nA > 0) {
            stitchPieces();
        }
    }

    // ReadAlign_multMapSelect.cpp:8-95 (outMultimapperOrder Old_2.4)
    void multMapSelect() {
        nTr = 0;
        if (nW == 0) return;
        maxScore = -10 * (int)Lread;
        for (uint iW = 0; iW < nW; iW++)
            if (maxScore < trAll[iW][0]->maxScore) maxScore = trAll[iW][0]->maxScore;
        if (maxScore != trBest->maxScore) { raise(STAR_EXIT_BUG, "BUG: maxScore!=trBest->maxScore in multMapSelect"); return; }
        trMult.clear();
        for (uint iW = 0; iW < nW; iW++) {
            for (uint iTr = 0; iTr < nWinTr[iW]; iTr++) {
                if ((trAll[iW][iTr]->maxScore + P.outFilterMultimapScoreRange) >= maxScore) {
                    trMult.push_back(trAll[iW][iTr]);
                    trMult[nTr]->Chr = trAll[iW][0]->Chr;
                    trMult[nTr]->Str = trAll[iW][0]->Str;
                    trMult[nTr]->roStr = trAll[iW][0]->roStr;
                    nTr++;
                }
            }
        }
        if (nTr > P.outFilterMultimapNmax || nTr == 0) return;
        for (uint iTr = 0; iTr < nTr; iTr++) {
            trMult[iTr]->roStart = trMult[iTr]->roStr == 0 ? trMult[iTr]->rStart : Lread - trMult[iTr]->rStart - trMult[iTr]->rLength;
            trMult[iTr]->cStart = trMult[iTr]->gStart - mapGen.v->chrStart[trMult[iTr]->Chr];
        }
        if (nTr == 1) {
            trMult[0]->primaryFlag = true;
        } else {
            int nbest = 0;
            if (P.outSAMmultNmax != (uint)-1) {
                for (uint itr = 0; itr < nTr; itr++) {
                    if (trMult[itr]->maxScore == maxScore) { std::swap(trMult[itr], trMult[nbest]); ++nbest; }
                }
            }
            if (P.outSAMprimaryFlagAllBestScore) {
                for (uint itr = 0; itr < nTr; itr++) if (trMult[itr]->maxScore == maxScore) trMult[itr]->primaryFlag = true;
            } else if (P.outSAMmultNmax != (uint)-1) {
                trMult[0]->primaryFlag = true;
            } else {
                trBest->primaryFlag = true;
            }
        }
    }

    // ReadAlign_mappedFilter.cpp:3-20
    void mappedFilter() {
        unmapType = -1;
        if (nW == 0) {
            unmapType = 0;
        } else if ((trBest->maxScore < P.outFilterScoreMin) || (trBest->maxScore < (intScore)(P.outFilterScoreMinOverLread * (Lread - 1)))
                   || (trBest->nMatch < P.outFilterMatchNmin) || (trBest->nMatch < (uint)(P.outFilterMatchNminOverLread * (Lread - 1)))) {
            unmapType = 1;
        } else if ((trBest->nMM > outFilterMismatchNmaxTotal) || (double(trBest->nMM) / double(trBest->rLength) > P.outFilterMismatchNoverLmax)) {
            unmapType = 2;
        } else if (nTr > P.outFilterMultimapNmax) {
            unmapType = 3;
        }
    }
};

void exportTr(const Tr& t, star_align_t* o) {
    memset(o, 0, sizeof(*o));
    for (uint i = 0; i < t.nExons; i++) {
        o->exG[i] = t.exons[i][EX_G]; o->exR[i] = (uint16_t)t.exons[i][EX_R]; o->exL[i] = (uint16_t)t.exons[i][EX_L];
        o->exFrag[i] = (uint8_t)t.exons[i][EX_iFrag];
        if (i + 1 < t.nExons) {
            o->canonSJ[i] = (int8_t)t.canonSJ[i]; o->sjAnnot[i] = t.sjAnnot[i]; o->sjStr[i] = t.sjStr[i];
            // shiftSJ is only defined for deletions/junctions found by the scan or the sjdb (stitchAlignToTranscript.cpp:26-27,243-244)
            bool hasShift = t.canonSJ[i] >= -1;
            o->shiftSJ[i][0] = hasShift ? (uint16_t)t.shiftSJ[i][0] : 0;
            o->shiftSJ[i][1] = hasShift ? (uint16_t)t.shiftSJ[i][1] : 0;
        }
    }
    o->nExons = (uint32_t)t.nExons; o->Chr = (uint32_t)t.Chr; o->Str = (uint8_t)t.Str; o->roStr = (uint8_t)t.roStr;
    o->primaryFlag = t.primaryFlag; o->sjMotifStrand = t.sjMotifStrand; o->iFrag = t.iFrag; o->maxScore = t.maxScore;
    o->nMatch = (uint32_t)t.nMatch; o->nMM = (uint32_t)t.nMM; o->nGap = (uint32_t)t.nGap; o->lGap = (uint32_t)t.lGap;
    o->nDel = (uint32_t)t.nDel; o->lDel = (uint32_t)t.lDel; o->nIns = (uint32_t)t.nIns; o->lIns = (uint32_t)t.lIns;
    o->nUnique = (uint32_t)t.nUnique; o->nAnchor = (uint32_t)t.nAnchor;
    o->rStart = (uint32_t)t.rStart; o->rLength = (uint32_t)t.rLength; o->roStart = (uint32_t)t.roStart;
    o->gStart = t.gStart; o->gLength = t.gLength; o->cStart = t.cStart;
}

struct ReadOut {
    star_read_result_t res;
    std::vector<star_align_t> aligns;
    std::vector<uint64_t> pc;       // nP x 8
    std::vector<uint64_t> wa;       // per window: header (iW, nWA, Str, Chr, gStart, gEnd) then nWA x 7
};

}  // namespace

struct star_oracle_ctx {
    Index index;
    star_params_t params;
    std::string lastError;
};

static thread_local std::string g_oracle_error;
static std::atomic<unsigned long long> g_karyChecked(0), g_karyMismatch(0), g_karyRounds(0), g_karyProbes(0);
static std::atomic<unsigned long long> g_winHistN[16], g_winHistNodes[16];

extern "C" {

int star_oracle_init(void** ctx, int /*device*/, const star_index_view_t* index, const star_params_t* params, uint32_t /*maxReads*/) {
    star_oracle_ctx* c = new star_oracle_ctx;
    c->index.init(index);
    c->params = *params;
    *ctx = c;
    return 0;
}

void star_oracle_destroy(void* ctx) { delete (star_oracle_ctx*)ctx; }
// k-ary seed search design check (STAR_ORACLE_KARY_CHECK=1): searches compared with the reference binary search / disagreements
void star_oracle_kary_stats(uint64_t* checked, uint64_t* mismatch) { *checked = g_karyChecked.load(); *mismatch = g_karyMismatch.load(); }
// dependent rounds (one round = probes issued together) and SA probes of the emulated k-ary searches (STAR_ORACLE_KARY_K = 2..32)
// analysis: windows and recursion nodes by number of seeds per window (bin 15 = 15 or more)
void star_oracle_window_hist(uint64_t* windows16, uint64_t* nodes16) { for (int q = 0; q < 16; q++) { windows16[q] = g_winHistN[q].load(); nodes16[q] = g_winHistNodes[q].load(); } }
void star_oracle_kary_cost(uint64_t* rounds, uint64_t* probes) { *rounds = g_karyRounds.load(); *probes = g_karyProbes.load(); }
const char* star_oracle_last_error(void) { return g_oracle_error.c_str(); }

static int oracle_run(star_oracle_ctx* c, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* stats,
                      star_oracle_dump_t* dump) {
    const uint32_t n = in->nReads;
    int nThreads = 1;
    if (const char* e = getenv("STAR_ORACLE_THREADS")) nThreads = std::max(1, atoi(e));
    nThreads = (int)std::min<uint32_t>(nThreads, std::max<uint32_t>(1, n / 64));
    std::vector<ReadOut> ro(n);
    std::vector<Counters> cnts(nThreads);
    std::vector<int> fatals(nThreads, 0);
    std::vector<std::string> fatalMsgs(nThreads);
    auto worker = [&](int t) {
        ReadAlign RA(c->params, c->index);
        uint32_t lo = (uint64_t)n * t / nThreads, hi = (uint64_t)n * (t + 1) / nThreads;
        for (uint32_t i = lo; i < hi; i++) {
            const uint64_t* off = in->seqOff + (uint64_t)i * in->nMates;
            const char* s0 = in->seq + off[0];
            uint l0 = off[1] - off[0];
            const char* s1 = in->nMates == 2 ? in->seq + off[1] : nullptr;
            uint l1 = in->nMates == 2 ? off[2] - off[1] : 0;
            ReadOut& o = ro[i];
            memset(&o.res, 0, sizeof(o.res));
            if (!RA.loadRead(s0, l0, s1, l1, in->nMates)) break;
            RA.mapOneRead();
            if (RA.fatal) break;
            if (dump) {
                o.pc.assign(RA.PCs.begin(), RA.PCs.begin() + RA.nP * PC_SIZE);
            }
            RA.multMapSelect();
            if (RA.fatal) break;
            RA.mappedFilter();
            o.res.unmapType = RA.unmapType;
            o.res.nTr = (uint32_t)RA.nTr;
            o.res.mapMarker = (uint32_t)RA.mapMarker;
            o.res.bestScore = RA.trBest->maxScore;
            o.res.bestNMM = (uint32_t)RA.trBest->nMM;
            o.res.bestRLength = (uint32_t)RA.trBest->rLength;
            o.res.Lread = (uint32_t)RA.Lread;
            o.res.bestTr = 0;
            if (RA.unmapType < 0) {
                o.res.nTrOut = (uint32_t)RA.nTr;
                o.aligns.resize(RA.nTr);
                for (uint k = 0; k < RA.nTr; k++) {
                    exportTr(*RA.trMult[k], &o.aligns[k]);
                    if (RA.trMult[k] == RA.trBest) o.res.bestTr = (uint32_t)k;
                }
            }
        }
        cnts[t] = RA.cnt;
        for (int q = 0; q < 16; q++) { g_winHistN[q] += RA.winHistN[q]; g_winHistNodes[q] += RA.winHistNodes[q]; }
        g_karyChecked += RA.karyChecked; g_karyMismatch += RA.karyMismatch; g_karyRounds += RA.karyRounds; g_karyProbes += RA.karyProbes;
        fatals[t] = RA.fatal;
        fatalMsgs[t] = RA.fatalMsg;
    };
    if (nThreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nThreads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < nThreads; t++) {
        if (fatals[t]) { g_oracle_error = fatalMsgs[t]; return fatals[t]; }
    }
    uint64_t nAl = 0;
    for (uint32_t i = 0; i < n; i++) {
        ro[i].res.trOffset = nAl;
        if (nAl + ro[i].aligns.size() > out->alignsCapacity) { g_oracle_error = "oracle: aligns capacity too small"; return STAR_EXIT_RUNTIME; }
        out->reads[i] = ro[i].res;
        for (auto& a : ro[i].aligns) out->aligns[nAl++] = a;
    }
    out->nAligns = nAl;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (auto& k : cnts) {
            stats->mmp_searches += k.searches; stats->mmp_sai_words += k.saiWords; stats->mmp_compare_calls += k.compareCalls;
            stats->mmp_bases_examined += k.basesExamined; stats->sa_enumerated += k.saEnum; stats->stitch_nodes += k.nodes; stats->stitch_leaves += k.leaves;
        }
    }
    if (dump) {
        uint64_t tot = 0;
        for (uint32_t i = 0; i < n; i++) tot += ro[i].pc.size();
        dump->pcOff = (uint64_t*)malloc((n + 1) * sizeof(uint64_t));
        dump->pc = (uint64_t*)malloc(std::max<uint64_t>(tot, 1) * sizeof(uint64_t));
        uint64_t p = 0;
        for (uint32_t i = 0; i < n; i++) {
            dump->pcOff[i] = p / PC_SIZE;
            if (!ro[i].pc.empty()) memcpy(dump->pc + p, ro[i].pc.data(), ro[i].pc.size() * sizeof(uint64_t));
            p += ro[i].pc.size();
        }
        dump->pcOff[n] = p / PC_SIZE;
    }
    return 0;
}

int star_oracle_map_chunk(void* ctx, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* stats) {
    return oracle_run((star_oracle_ctx*)ctx, in, out, stats, nullptr);
}

int star_oracle_map_chunk_dump(void* ctx, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* stats, star_oracle_dump_t* dump) {
    return oracle_run((star_oracle_ctx*)ctx, in, out, stats, dump);
}

void star_oracle_dump_free(star_oracle_dump_t* d) {
    free(d->pcOff); free(d->pc);
    d->pcOff = nullptr; d->pc = nullptr;
}

int star_oracle_set_sj_novel(void* ctx, const uint64_t* sjStart, const uint64_t* sjEnd, uint64_t n) {
    star_oracle_ctx* c = (star_oracle_ctx*)ctx;
    c->index.sjNovelOn = true;
    c->index.sjNovelStart.assign(sjStart, sjStart + n);
    c->index.sjNovelEnd.assign(sjEnd, sjEnd + n);
    return 0;
}

// ---- junction insertion: CPU restatement of the two device steps (checker for star_gpu_sjdb_*) -------------------------------------
// Sequential, in the reference's own shape: suffixArraySearch1 / compareSeqToGenome1 / compareRefEnds (SuffixArrayFuns.cpp:221-351)
// per suffix, and the single two-pointer sweep with PackedArray::writePacked of sjdbBuildIndex.cpp:141-214.
namespace {
struct SjdbOracle {
    const star_index_view_t* v;
    uint64_t saGet(uint64_t i) const {   // PackedArray.h:24-32
        const uint64_t b = i * (v->GstrandBit + 1);
        uint64_t w;
        memcpy(&w, v->SA + b / 8, 8);
        return (w >> (b % 8)) & (~0ULL >> (64 - (v->GstrandBit + 1)));
    }
    // compareSeqToGenome1 with dirR = true, gInsert = -1: s0 = the insert text, s1 = its complement (as the reference passes both)
    uint64_t compare1(const uint8_t* s0, const uint8_t* s1, uint64_t S, uint64_t N, uint64_t L, uint64_t iSA, int& compRes) const {
        uint64_t SAstr = saGet(iSA);
        const bool dirG = (SAstr >> v->GstrandBit) == 0;
        SAstr &= ~(1ULL << v->GstrandBit);
        if (dirG) {
            const uint8_t* s = s0 + S + L;
            const uint8_t* g = v->G + SAstr + L;
            for (uint64_t ii = 0; ii < N - L; ii++) {
                if (s[ii] != g[ii]) { compRes = s[ii] > g[ii] ? 1 : -1; return ii + L; }
                else if (s[ii] == 5) { compRes = 1; /* compareRefEnds: strG, strR, SAstr < -1 */ return ii + L; }
            }
            return N;
        } else {
            const uint8_t* s = s1 + S + L;
            const uint8_t* g = v->G + v->nGenome - 1 - SAstr - L;
            for (uint64_t ii = 0; ii < N - L; ii++) {
                if (s[ii] != *(g - ii)) {
                    uint8_t a = s[ii], b = *(g - ii);
                    if (a < 4) a = 3 - a;
                    if (b < 4) b = 3 - b;
                    compRes = a > b ? 1 : -1;
                    return ii + L;
                } else if (s[ii] == 5) { compRes = -1; /* compareRefEnds: !strG, strR */ return ii + L; }
            }
            return N;
        }
    }
    uint64_t search1(const uint8_t* s0, const uint8_t* s1, uint64_t S) const {
        const uint64_t N = 10000;
        int compRes = 0;
        uint64_t i1 = 0, i2 = v->nSA - 1;
        uint64_t L1 = compare1(s0, s1, S, N, 0, i1, compRes);
        if (compRes < 0) return 0;
        uint64_t L2 = compare1(s0, s1, S, N, 0, i2, compRes);
        if (compRes > 0) return (uint64_t)-2;
        uint64_t L = std::min(L1, L2);
        while (i1 + 1 < i2) {
            const uint64_t i3 = i1 / 2 + i2 / 2 + (i1 % 2 + i2 % 2) / 2;
            const uint64_t L3 = compare1(s0, s1, S, N, L, i3, compRes);
            if (L3 == N) return i3;
            if (compRes > 0) { i1 = i3; L1 = L3; } else if (compRes < 0) { i2 = i3; L2 = L3; }
            L = std::min(L1, L2);
        }
        return i2;
    }
};
}  // namespace

int star_oracle_sjdb_open(void** h, int, const star_index_view_t* oldIndex) {
    SjdbOracle* o = new SjdbOracle;
    o->v = oldIndex;
    *h = o;
    return 0;
}
int star_oracle_sjdb_search(void* h, const uint8_t* Gsj, uint64_t sjdbN, uint64_t sjdbLength, const uint8_t* skipSeq, uint64_t* indArray) {
    const SjdbOracle* o = (const SjdbOracle*)h;
    const uint64_t nText = 2 * sjdbN * sjdbLength + 1;
    std::vector<uint8_t> G1c(nText + 16, 5);
    for (uint64_t i = 0; i < nText; i++) G1c[i] = Gsj[i] < 4 ? 3 - Gsj[i] : Gsj[i];   // complementSeqNumbers
    for (uint64_t isj = 0; isj < 2 * sjdbN; isj++)
        for (uint64_t istart = 0; istart < sjdbLength; istart++) {
            const uint64_t k = isj * sjdbLength + istart;
            if (skipSeq[isj] || Gsj[k] > 3) indArray[2 * k] = (uint64_t)-1;
            else indArray[2 * k] = o->search1(Gsj + isj * sjdbLength, G1c.data() + isj * sjdbLength, istart);
            indArray[2 * k + 1] = k;
        }
    return 0;
}
int star_oracle_sjdb_merge_sa(void* h, const uint64_t* ind, uint64_t nInd, uint64_t nGsj, uint64_t nGsjNew, uint64_t sjdbLength, const uint32_t* oldSJind,
                              uint8_t* SAnew, uint64_t nSAnewByte) {
    const SjdbOracle* o = (const SjdbOracle*)h;
    const star_index_view_t* v = o->v;
    const uint32_t bits = v->GstrandBit + 1;
    const uint64_t mask = ~0ULL >> (64 - bits);
    std::vector<uint8_t> buf(nSAnewByte + 16, 0);
    auto put = [&](uint64_t jj, uint64_t x) {   // PackedArray::writePacked
        const uint64_t b = jj * bits, S = b % 8;
        uint64_t w;
        memcpy(&w, buf.data() + b / 8, 8);
        w = (w & ~(mask << S)) | (x << S);
        memcpy(buf.data() + b / 8, &w, 8);
    };
    const uint64_t N2bit = 1ULL << v->GstrandBit, strandMask = ~N2bit;
    const uint64_t sjG = v->chrStart[v->nChrReal], nGenome1 = v->nGenome, nGenome = sjG + nGsj;
    uint64_t isj = 0, isa2 = 0;
    for (uint64_t isa = 0; isa < v->nSA; isa++) {
        while (isj < nInd && isa == ind[isj * 2]) {
            uint64_t ind1 = ind[isj * 2 + 1];
            if (ind1 < nGsj) ind1 += sjG; else ind1 = (ind1 - nGsj) | N2bit;
            put(isa2, ind1);
            ++isa2; ++isj;
        }
        uint64_t ind1 = o->saGet(isa);
        if ((ind1 & N2bit) > 0) {
            uint64_t ind1s = nGenome1 - (ind1 & strandMask);
            if (ind1s >= sjG) {
                const uint64_t sj1 = (ind1s - sjG) / sjdbLength;
                if (sj1 < v->sjdbN) ind1s += ((uint64_t)oldSJind[sj1] - sj1) * sjdbLength;
                ind1 = (nGenome - ind1s) | N2bit;
            } else ind1 += nGsjNew;
        } else if (ind1 >= sjG) {
            const uint64_t sj1 = (ind1 - sjG) / sjdbLength;
            if (sj1 < v->sjdbN) ind1 += ((uint64_t)oldSJind[sj1] - sj1) * sjdbLength;
        }
        put(isa2, ind1);
        ++isa2;
    }
    for (; isj < nInd; isj++) {
        uint64_t ind1 = ind[isj * 2 + 1];
        if (ind1 < nGsj) ind1 += sjG; else ind1 = (ind1 - nGsj) | N2bit;
        put(isa2, ind1);
        ++isa2;
    }
    memcpy(SAnew, buf.data(), nSAnewByte);
    return 0;
}
void star_oracle_sjdb_close(void* h) { delete (SjdbOracle*)h; }

// ---- index generation: CPU restatement of the suffix sort (checker for star_gpu_sa_build) -------------------------------------------
// std::sort with the comparison of funCompareSuffixes (Genome_genomeGenerate.cpp:29-89) written character by character on the
// text G + reverse complement: first difference decides (codes compare by value, 5 largest); a 5 at the same offset in both ends the
// comparison and the smaller text position sorts first.
int star_oracle_sa_build(int, const uint8_t* G, uint64_t nGenome, uint32_t GstrandBit, uint64_t nSA, uint8_t* SA, uint64_t nSAbyte) {
    const uint64_t n = 2 * nGenome;
    std::vector<uint8_t> T(n + 256, 5);
    for (uint64_t i = 0; i < nGenome; i++) { T[i] = G[i]; T[n - 1 - i] = G[i] < 4 ? 3 - G[i] : G[i]; }
    std::vector<uint64_t> pos;
    pos.reserve(nSA);
    for (uint64_t i = 0; i < n; i++) if (T[i] < 4) pos.push_back(i);
    if (pos.size() != nSA) return STAR_EXIT_BUG;
    const uint8_t* t = T.data();
    std::sort(pos.begin(), pos.end(), [t](uint64_t a, uint64_t b) {
        for (uint64_t k = 0;; k++) {
            const uint8_t ca = t[a + k], cb = t[b + k];
            if (ca != cb) return ca < cb;
            if (ca == 5) return a < b;
        }
    });
    const uint32_t bits = GstrandBit + 1;
    const uint64_t N2bit = 1ULL << GstrandBit;
    std::vector<uint8_t> buf(nSAbyte + 16, 0);
    for (uint64_t j = 0; j < nSA; j++) {
        const uint64_t x = pos[j] < nGenome ? pos[j] : ((pos[j] - nGenome) | N2bit);
        const uint64_t b = j * bits, S = b % 8;
        uint64_t w;
        memcpy(&w, buf.data() + b / 8, 8);
        w |= x << S;
        memcpy(buf.data() + b / 8, &w, 8);
    }
    memcpy(SA, buf.data(), nSAbyte);
    return 0;
}

static const star_engine_vtbl_t g_oracle_vtbl = {star_oracle_init, star_oracle_map_chunk, star_oracle_destroy, star_oracle_last_error,
                                                 star_oracle_sjdb_open, star_oracle_sjdb_search, star_oracle_sjdb_merge_sa, star_oracle_sjdb_close, star_oracle_sa_build, star_oracle_set_sj_novel};
const star_engine_vtbl_t* star_oracle_engine(void) { return &g_oracle_vtbl; }

}  // extern "C"
