// dev.cuh — device-side data layout of the star-b200 engine (sm_100a).
//
// Everything the kernels read is resident in HBM for the lifetime of the context:
//   G    genome, 1 byte / base, codes 0..5, 256 bytes of code 5 on both sides  (reference Genome.h:26, F2 of SURVEY.md)
//   SA   suffix array, (GstrandBit+1)-bit packed, read as aligned 64-bit words  (reference PackedArray.h:24-32)
//   SAi  prefix table, (GstrandBit+3)-bit packed
//   chrBin, chrStart, chrLength, sjdb arrays (SoA)
// Per-read scratch lives in per-lane arenas (persistent lanes, see stitch.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/star_b200.h"
#include "warp_prims.cuh"

namespace starb {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

struct DevIndex {
    const u8* G;            // points at base 0 (padding before/after is readable)
    u64 nGenome;
    const u64* SA;          // packed words
    u64 nSA;
    const u64* SAi;
    u64 nSAi;
    u32 GstrandBit, saBits, saiBits, gSAindexNbases, gChrBinNbits, nChrReal;
    u64 GstrandMask, SAiMarkAbsentMaskC, SAiMarkNmaskC, SAiMarkNmask;
    u64 genomeSAindexStart[20];
    const u32* chrBin;
    u64 chrBinN;
    const u64* chrStart;
    const u64* chrLength;
    u64 sjdbN, sjdbOverhang, sjdbLength, sjGstart;
    const u64* sjdbStart;
    const u64* sjdbEnd;
    const u64* sjDstart;
    const u64* sjAstart;
    const u8* sjdbMotif;
    const u8* sjdbShiftLeft;
    const u8* sjdbShiftRight;
    const u8* sjdbStrand;
    // 2nd stage of --outFilterType BySJout: novel junctions that passed the filters, sorted by (start, end); sjNovelOn = 0: no filtering
    const u64* sjNovelStart;
    const u64* sjNovelEnd;
    u64 sjNovelN;
    u32 sjNovelOn;
    // integer thresholds of the genomic-length score: value v[k] applies for gLen >= thr[k] (host libm, see engine_api.cu)
    const u64* log2Thr;
    const int* log2Val;
    int log2N;
};

// PackedArray::operator[] (reference PackedArray.h:24-32) on 64-bit aligned words: the reference does one
// unaligned 8-byte load at byte b/8 and shifts by b%8; two aligned words give the same bits.
__device__ __forceinline__ u64 packedGet(const u64* __restrict__ w, u32 bits, u64 ii) {
    u64 b = ii * bits;
    u64 wi = b >> 6;
    u32 sh = (u32)(b & 63);
    u64 lo = SB_LDG(w + wi);
    u64 v = lo >> sh;
    if (sh + bits > 64) {
        u64 hi = SB_LDG(w + wi + 1);
        v |= hi << (64 - sh);
    }
    return v & ((1ULL << bits) - 1ULL);
}

// one stored seed ("piece"), reference PC[][8] (IncludeDefine.h:181-189); SAend = SAstart + Nrep - 1
struct Piece {
    u64 SAstart;
    u16 rStart, Length;
    u16 Nrep;
    u8 Dir, iFrag;
};

// per-read header produced by the prep kernel
struct ReadInfo {
    u32 Lread;
    u16 readLength[2];
    u16 nP;            // number of stored pieces after seeding
    u32 nA;            // total number of loci of stored pieces
    u32 mapMarker;
    u32 multNminL;
    u16 Nsplit;
    u16 split1_0;      // splitR[1][0] when Nsplit==0 (min good piece length)
    u32 outFilterMismatchNmaxTotal;
    u32 flags;         // bit0: overflowed a fast-path cap (needs the slow path); bit1: fatal "too many pieces"
    // algorithmic work of THIS read (overwritten when the read is redone on the slow path, so nothing is counted twice)
    u32 cSearches, cSaiWords, cCompare, cBases, cSaEnum, cNodes, cLeaves, cSlow;
};

struct WorkCounters {  // algorithmic work counters, accumulated per kernel launch
    u64 searches, saiWords, compareCalls, basesExamined, saEnum, nodes, leaves, slowReads;
};

// seed inside a window, reference WA[][7] (IncludeDefine.h:197-204)
struct Seed {
    u64 gStart;
    u32 sjA;
    u16 rStart, Length;
    u16 Nrep;
    u8 Anchor, iFrag;
};

struct Window {  // reference WC[][4] + nWA/WALrec
    u32 gStart, gEnd;   // bins; dead window: gStart=1,gEnd=0
    u32 Chr;
    u16 nWA;
    u16 WALrec;
    u8 Str;
    u8 pad[3];
};

struct Exon {
    u64 G;
    u32 sjA;
    u16 R, L;
    u16 shL, shR;       // shiftSJ of the junction AFTER this exon
    u8 iFrag;
    signed char canon;  // canonSJ of the junction after this exon
    u8 annot, sjStr;
};

struct TrHead {  // Transcript.h scalars used on the path
    u64 gStart, gLength;
    u32 rStart, rLength;
    int maxScore;
    u32 nMatch, nMM, mappedLength;
    u32 nGap, lGap, nDel, lDel, nIns, lIns;
    u16 nUnique, nAnchor;
    u16 nExons;
    signed char iFrag;
    u8 sjMotifStrand;
    u8 primaryFlag;
    u8 pad[3];
};

struct DevTr {
    TrHead h;
    Exon ex[STAR_MAX_N_EXONS];
};

// capacities of one lane's arena: the fast path uses small ones, the slow path the reference's own limits
struct Caps {
    u32 maxP;        // pieces per read (slab of the seed kernel)
    u32 maxW;        // windows per read
    u32 maxTr;       // transcript pool per read
    u32 spw;         // seedPerWindowNmax
    u32 nOut;        // staged alignments per read (outFilterMultimapNmax)
    u32 binFilter = 0;   // setup kernel: hashed (strand, bin) bitmap in front of the window lookup of a locus (STAR_B200_BIN_FILTER=1; measured: no gain)
    u32 sortMinW = 12;   // reads with more windows than this look a locus' window up by bisection (sorted index) instead of a scan
    u64 arenaBytes;
};

}  // namespace starb
