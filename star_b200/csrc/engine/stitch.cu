// stitch.cu — seed -> window assignment, seed stitching, end extension, multimapper selection (sm_100a).
//
// What is computed (bit-exact with the reference, proven by tests/ against oracle/):
//   ReadAlign::stitchPieces                 reference source/ReadAlign_stitchPieces.cpp:12-350
//   createExtendWindowsWithAlign            source/ReadAlign_createExtendWindowsWithAlign.cpp:7-84
//   assignAlignToWindow, sjAlignSplit       source/ReadAlign_assignAlignToWindow.cpp:6-130, sjAlignSplit.cpp:3-15
//   stitchWindowAligns (short-read stitcher) source/stitchWindowAligns.cpp:8-353
//   stitchAlignToTranscript, binarySearch2   source/stitchAlignToTranscript.cpp:9-415, binarySearch2.cpp:3-43
//   extendAlign, blocksOverlap               source/extendAlign.cpp:6-92, blocksOverlap.cpp:3-40
//   multMapSelect, mappedFilter              source/ReadAlign_multMapSelect.cpp:8-95, ReadAlign_mappedFilter.cpp:3-20
//
// How (B200-first, NOT the reference's structure):
//  * three execution shapes share the device functions of this file:
//      - the flattened path of stitch_flat.cuh (default for every read with >= 4 loci): setup (one warp per read) ->
//        sub-tree tasks (one warp per task, all lanes on the same scalar path, byte loops 32 positions per step: the COOP
//        template variants below) -> ordered recording (one warp per read);
//      - stitch_kernel: one read per lane (persistent lanes, lockstep state machine) for reads with very few loci and as
//        the engine of the overflow tiers, with stitch_heavy_kernel (one warp per read) for the DFS-heavy reads of a tier;
//    every shape uses private arenas in HBM (windows, seeds, transcript pool) with caps and overflow flags;
//  * the dense per-read winBin[2][nGenome>>16] array (190 KB memset per read in the reference, F8) is replaced by the
//    window interval list itself: bins owned by a window are exactly [gStart,gEnd] of a live window on that strand,
//    so "is the bin owned / nearest owned bin within winAnchorDistNbins" are interval queries (DESIGN.md proves the
//    equivalence, incl. the flank extension and the kill-right-window case);
//  * the 2^n include/exclude recursion that copies a 1.7 kB Transcript per node (30 % of the reference's CPU time) is an
//    explicit DFS over ONE shared transcript with a per-level undo record (head + last exon, 104 bytes): stitching only
//    ever appends an exon and edits the last one.  Leaves materialise a private copy, so visiting order, dedup and
//    tie-breaks are exactly the reference's;
//  * log2() of the genomic-length score is an integer threshold table computed by the host libm (CUDA's log2 is not
//    correctly rounded); every other floating-point expression is a single IEEE multiply/divide and is evaluated
//    in fp64 as written in the reference.
#include "dev.cuh"
#include "stitch_types.cuh"

// Warp-shared transcript of the cooperative kernels (ONE transcript per warp in shared memory, every lane executes the same scalar path).
// SB_PRIVATE_COPIES = 1: stitchAlignToTranscript / evalLeaf work on private copies of the head and the touched exons and lane 0 writes them
// back (strict single writer; what the host emulation runs, where the 32 lanes are free-running threads).  0 (the GPU build): the lanes of
// the converged warp update the shared transcript in place — every lane executes the same store instruction with the same value, which the
// hardware issues as one shared-memory transaction; no lane branches on lane-dependent data between a store and the next load, and every
// cooperative section ends in a __syncwarp.  Measured on B200: the private copies cost +29 % on the task kernel (80 + 2 x 24 bytes per lane
// per stitch, 128 registers + 128 B of spills); -DSTAR_B200_PRIVATE_COPIES=1 builds the strict variant for the GPU (profiles/r02_summary.md).
#if defined(STAR_CUDA_HOST_SHIM) || defined(STAR_B200_PRIVATE_COPIES)
#define SB_PRIVATE_COPIES 1
#else
#define SB_PRIVATE_COPIES 0
#endif

namespace starb {

#define SJA_NONE 0xFFFFFFFFu

// optional cycle accounting (per warp, lane 0 adds at kernel end): light kernel phases [0..8], heavy kernel [16..]
__device__ unsigned long long g_prof[32];
#define PROF_ADD(slot, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_prof[slot], (unsigned long long)(v)); } while (0)

#define HEAVY_SPLIT_MIN 6        // heavy kernel: windows with more seeds than this are cut into 2^(nWA-6) (max 256) prefix sub-trees
#define LOCI_PER_STEP 1          // SA loci handled per lockstep step (their SA words are loaded back to back, latency overlapped)
#define STAR_DFS_MAX_DEPTH 52   // seedPerWindowNmax (<=50 on the local-memory fast build) + 2

// Undo record of one successful (or attempted) seed include: the transcript head and the last exon before the include, and the
// DFS cursor (Score, tR2, tG2).  Only include levels need one (<= MAX_N_EXONS+1 deep); exclude levels change nothing.
struct Frame {
    TrHead h;
    Exon last;
    u64 tG2;
    u32 tR2;
    int Score;
    u32 pad[2];
};

static_assert(sizeof(Frame) == 128, "arenaSize() in engine_api.cu assumes 128-byte frames");
#define STAR_UNDO_DEPTH (STAR_MAX_N_EXONS + 2)

struct Lane {
    const DevIndex* ix;
    const star_params_t* P;
    const u8* R0;   // Read1[0] in shared memory
    const u8* R2;   // Read1[2] (reverse complement)
    const u8* R;    // orientation of the current window
    u32 Lread;
    u16 readLength[2];
    u32 outFilterMismatchNmaxTotal;
    int maxScoreMate[2];
    // arena
    Window* win;
    Seed* wa;
    u16* trPtr;
    DevTr* pool;
    u16* winBase;   // per non-empty window iW1: first index into trPtr
    u16* winN;      // per iW1: nWinTr
    Frame* stack;
    DevTr* cur;
    DevTr* leaf;
    u32 nW;
    Caps caps;
    u32 overflow;
    u64 saEnum, nodes, leaves;
    u8* ph;              // per DFS level: 0 new, 1 include branch taken, 2 exclude branch taken, 3 forced include
    int level;           // current DFS level (= seed index iA); -1 when the window is finished
    int nInc;            // number of live undo records
    int Score;           // DFS cursor
    u32 tR2;
    u64 tG2;
    int leafScore;       // pending leaf (set by dfsStep)
    u32 leafR2;
    u64 leafG2;
    u64 inclMask;        // bit k set <=> seed k of the window is included on the current DFS path
    u32 memoNMM; bool memoMotifOk;   // set by stitchAlignToTranscript: mismatches of this stitch / junction-motif rule
    struct StitchMemo* memo;         // heavy kernel: per-warp table of same-fragment stitch results (NULL = off)
    u32 memoMask;                    // slots-1
    u64 memoBase;                    // (epoch, window) part of the key
    int lastSeed;                    // index of the last included seed on the current path (-1 = none)
    u64 memoHit, memoMiss;
    u32 coop;            // 1: all 32 lanes of the warp run this Lane with identical state (flat_dfs_warp_kernel): byte loops are cooperative
    u32 forceDepth;      // heavy path: the first forceDepth include/exclude decisions are fixed (prefix sub-tree task)
    u32 forceBits;       // bit (forceDepth-1-k) set <=> seed k is EXCLUDED (so that ascending task id = DFS order, include first)
    // per-read stitching state (ReadAlign_stitchPieces.cpp:262-350)
    u32 trNtotal, nW1;
    int bestPool, bestScore;
    u64 bestGLength;
};

__device__ __forceinline__ u8 Gat(const Lane& ln, u64 pos) { return __ldg(ln.ix->G + (i64)pos); }

// binarySearch2.cpp:3-43
__device__ int binarySearch2(u64 x, u64 y, const u64* __restrict__ X, const u64* __restrict__ Y, int N) {
    if (N == 0 || x > X[N - 1] || x < X[0]) return -1;
    int i1 = 0, i2 = N - 1, i3 = N / 2;
    #pragma unroll 1
    while (i2 > i1 + 1) {
        i3 = (i1 + i2) / 2;
        if (X[i3] > x) i2 = i3; else i1 = i3;
    }
    if (x == X[i1]) i3 = i1;
    else if (x == X[i2]) i3 = i2;
    else return -1;
    #pragma unroll 1
    for (int jj = i3; jj >= 0; jj--) {
        if (x != X[jj]) break;
        else if (y == Y[jj]) return jj;
    }
    #pragma unroll 1
    for (int jj = i3; jj < N; jj++) {
        if (x != X[jj]) return -1;
        else if (y == Y[jj]) return jj;
    }
    return -2;
}

struct ExtRes { u32 extendL; int maxScore; u32 nMatch, nMM; };

// extendAlign.cpp:6-92 (local extension, extendToEnd==false) executed by a whole warp: all 32 lanes call with IDENTICAL arguments and
// get identical results; lane j examines base 32*c+j of chunk c.  Equivalence with the sequential loop:
//  * the loop ends at the first base that is a stop (genome edge / padding, mate spacer, i==L) or at the first mismatch whose
//    running mismatch count already reaches capEnd — both are "first lane with the property" (ballot + ffs);
//  * Score/nMatch/nMM at a base are prefix counts of matches / mismatches (popc of ballot masks below the lane);
//  * a base is recorded when it is a match, its cap test holds and its Score exceeds every previously recorded Score, so the final
//    record is the FIRST base that attains the maximum Score among the matches passing the cap test (if that maximum is > 0).
__device__ __forceinline__ bool coopExtendAlign(const u8* R, const u8* G, const u64 nG, const u64 Lread, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev, u64 nMMmax, double pMMmax, ExtRes& res) {
    const u32 lane = threadIdx.x & 31;
    const u32 below = (1u << lane) - 1;
    res.maxScore = 0;
    double capEnd = pMMmax * double(Lprev + L);
    const double nMMmaxD = double(nMMmax);
    if (nMMmaxD < capEnd) capEnd = nMMmaxD;
    int Score = 0, nMatch = 0, nMM = 0;   // state after the bases of the previous chunks
    #pragma unroll 1
    for (u64 base = 0; base < L; base += 32) {
        const u64 i = base + lane;
        bool stop = i >= L;
        u8 g = 0, r = 0;
        if (!stop) {
            const u64 gpos = gStart + (u64)((i64)dG * (i64)i);
            const u64 rpos = rStart + (u64)((i64)dR * (i64)i);
            if (gpos >= nG || rpos >= Lread) stop = true;   // (u64)-1 = before the genome start; beyond the end the padding (5) stops the loop
            else {
                g = __ldg(G + gpos); r = R[rpos];
                if (g == 5 || r == STAR_MARK_FRAG_SPACER_BASE) stop = true;
            }
        }
        const u32 stopMask = __ballot_sync(0xffffffffu, stop);
        u32 nValid = stopMask ? (u32)__ffs(stopMask) - 1 : 32;
        bool ended = stopMask != 0;
        const bool notN = !(r > 3 || g > 3);
        const bool isX = lane < nValid && notN && g != r;
        u32 xMask = __ballot_sync(0xffffffffu, isX);
        const bool brk = isX && double((u64)(nMM + __popc(xMask & below)) + nMMprev) >= capEnd;
        const u32 brkMask = __ballot_sync(0xffffffffu, brk);
        if (brkMask) { nValid = (u32)__ffs(brkMask) - 1; ended = true; }
        const u32 vm = nValid >= 32 ? 0xffffffffu : ((1u << nValid) - 1);
        const u32 mMask = __ballot_sync(0xffffffffu, notN && g == r) & vm;
        xMask &= vm;
        const int mIncl = __popc(mMask & (below | (1u << lane)));
        const int xB = __popc(xMask & below);
        const int Score_i = Score + mIncl - xB, nMM_i = nMM + xB, nMatch_i = nMatch + mIncl;
        bool cand = (mMask >> lane) & 1u;
        if (cand) {
            double cap = pMMmax * double(Lprev + i + 1);
            if (nMMmaxD < cap) cap = nMMmaxD;
            cand = double((u64)nMM_i + nMMprev) <= cap;
        }
        const int best = __reduce_max_sync(0xffffffffu, cand ? Score_i : (int)0x80000000);
        if (best > res.maxScore) {
            const u32 src = (u32)__ffs(__ballot_sync(0xffffffffu, cand && Score_i == best)) - 1;
            res.maxScore = best;
            res.extendL = (u32)base + src + 1;
            res.nMatch = (u32)__shfl_sync(0xffffffffu, nMatch_i, src);
            res.nMM = (u32)__shfl_sync(0xffffffffu, nMM_i, src);
        }
        Score += __popc(mMask) - __popc(xMask); nMatch += __popc(mMask); nMM += __popc(xMask);
        if (ended) break;
    }
    return res.extendL > 0;
}

// extendAlign.cpp:6-92.  R/G are addressed through the lane (R orientation already selected).
// COOP: the caller is a warp executing one Lane uniformly (flat_dfs_warp_kernel).
struct ExtOut { ExtRes r; bool ok; };
// One copy per kernel (four call sites).  Takes the sequence pointers instead of the Lane so that the caller's Lane never escapes
// (its fields can then live in registers).
template <bool COOP>
__device__ __noinline__ ExtOut extendAlignShared(const u8* Rbase, const u8* Gbase, u64 nG, u64 Lread, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev,
                                                 u64 nMMmax, double pMMmax, bool extendToEnd);
template <bool COOP = false>
__device__ __forceinline__ bool extendAlign(const Lane& ln, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev, u64 nMMmax, double pMMmax,
                                            bool extendToEnd, ExtRes& res) {
    const ExtOut o = extendAlignShared<COOP>(ln.R, ln.ix->G, ln.ix->nGenome, ln.Lread, rStart, gStart, dR, dG, L, Lprev, nMMprev, nMMmax, pMMmax, extendToEnd);
    res = o.r;
    return o.ok;
}
template <bool COOP>
__device__ bool extendAlignBody(const u8* Rbase, const u8* Gbase, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev, u64 nMMmax, double pMMmax,
                                bool extendToEnd, ExtRes& res);
template <bool COOP>
__device__ __noinline__ ExtOut extendAlignShared(const u8* Rbase, const u8* Gbase, u64 nG, u64 Lread, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev,
                                                 u64 nMMmax, double pMMmax, bool extendToEnd) {
    ExtOut o;
    o.r.extendL = 0; o.r.maxScore = 0; o.r.nMatch = 0; o.r.nMM = 0;   // (every caller passes a zeroed result)
    if constexpr (COOP) {
        if (!extendToEnd) { o.ok = coopExtendAlign(Rbase, Gbase, nG, Lread, rStart, gStart, dR, dG, L, Lprev, nMMprev, nMMmax, pMMmax, o.r); return o; }
    }
    o.ok = extendAlignBody<COOP>(Rbase, Gbase, rStart, gStart, dR, dG, L, Lprev, nMMprev, nMMmax, pMMmax, extendToEnd, o.r);
    return o;
}
// (COOP: only reached for extendToEnd, a cold path there: kept out of line)
template <bool COOP>
__device__ bool extendAlignBody(const u8* Rbase, const u8* Gbase, u64 rStart, u64 gStart, int dR, int dG, u64 L, u64 Lprev, u64 nMMprev, u64 nMMmax, double pMMmax,
                                bool extendToEnd, ExtRes& res) {
    int Score = 0, nMatch = 0, nMM = 0;
    res.maxScore = 0;
    const u8* R = Rbase + rStart;
    const u8* G = Gbase + (i64)gStart;
    if (extendToEnd) {
        int iExt;
        #pragma unroll 1
        for (iExt = 0; iExt < (int)L; iExt++) {
            int iS = dR * iExt, iG = dG * iExt;
            u8 g;
            if ((gStart + (u64)(i64)iG) == (u64)(-1LL) || (g = __ldg(G + iG)) == 5) {
                res.extendL = 0; res.maxScore = -999999999; res.nMatch = 0; res.nMM = (u32)(nMMmax + 1);
                return true;
            }
            u8 r = R[iS];
            if (r == STAR_MARK_FRAG_SPACER_BASE) break;
            if (r > 3 || g > 3) continue;
            if (g == r) { nMatch++; Score += 1; } else { nMM++; Score -= 1; }
        }
        if (iExt > 0) {
            res.extendL = (u32)iExt; res.maxScore = Score; res.nMatch = (u32)nMatch; res.nMM = (u32)nMM;
            return true;
        }
        return false;
    }
    double capEnd = pMMmax * double(Lprev + L);
    double nMMmaxD = double(nMMmax);
    if (nMMmaxD < capEnd) capEnd = nMMmaxD;
    #pragma unroll 1
    for (int i = 0; i < (int)L; i++) {
        int iS = dR * i, iG = dG * i;
        if ((gStart + (u64)(i64)iG) == (u64)(-1LL)) break;
        u8 g = __ldg(G + iG);
        u8 r = R[iS];
        if (g == 5 || r == STAR_MARK_FRAG_SPACER_BASE) break;
        if (r > 3 || g > 3) continue;
        if (g == r) {
            nMatch++; Score += 1;
            if (Score > res.maxScore) {
                double cap = pMMmax * double(Lprev + (u64)i + 1);
                if (nMMmaxD < cap) cap = nMMmaxD;
                if (double((u64)nMM + nMMprev) <= cap) {
                    res.extendL = (u32)(i + 1); res.maxScore = Score; res.nMatch = (u32)nMatch; res.nMM = (u32)nMM;
                }
            }
        } else {
            if (double((u64)nMM + nMMprev) >= capEnd) break;
            nMM++; Score -= 1;
        }
    }
    return res.extendL > 0;
}

// stitchAlignToTranscript.cpp:9-415 operating on the shared DFS transcript `t`
// COOP: executed by a whole warp with identical arguments; the byte loops run 32 positions per step (same results, see the
// comments at each loop).
template <bool COOP = false>
__device__ int stitchAlignToTranscript(Lane& ln, u64 rAend, u64 gAend, u64 rBstart, u64 gBstart, u64 L, u32 iFragB, u32 sjAB, DevTr* t) {
    const star_params_t& P = *ln.P;
    const DevIndex& g = *ln.ix;
    if (t->h.nExons >= STAR_MAX_N_EXONS) return -1000010;
    // COOP (the transcript is ONE copy per warp in shared memory): every lane works on private copies of the head, the last exon and the
    // new exon and lane 0 alone writes them back on success — no lane ever reads shared state another lane is writing.
    const u32 nEx0 = t->h.nExons;
    constexpr bool PRIV = COOP && SB_PRIVATE_COPIES;
    TrHead hL; Exon eAL, eBL;
    if constexpr (PRIV) { hL = t->h; eAL = t->ex[nEx0 - 1]; eBL = t->ex[nEx0]; }
    TrHead& h = PRIV ? hL : t->h;
    const u8* R = ln.R;
    int Score = 0;
    Exon& eA = PRIV ? eAL : t->ex[nEx0 - 1];
    Exon& eB = PRIV ? eBL : t->ex[nEx0];
    const u64 outFilterMismatchNmaxTotal = ln.outFilterMismatchNmaxTotal;

    if (__builtin_expect(sjAB != SJA_NONE && eA.sjA == sjAB && eA.iFrag == iFragB && rBstart == rAend + 1 && gAend + 1 < gBstart, 0)) {
        if (g.sjdbMotif[sjAB] == 0 && (L <= g.sjdbShiftRight[sjAB] || eA.L <= g.sjdbShiftLeft[sjAB])) return -1000006;
        eB.L = (u16)L; eB.R = (u16)rBstart; eB.G = gBstart;
        eA.canon = (signed char)g.sjdbMotif[sjAB];
        eA.shL = g.sjdbShiftLeft[sjAB]; eA.shR = g.sjdbShiftRight[sjAB];
        eA.annot = 1;
        eA.sjStr = g.sjdbStrand[sjAB];
        h.nExons++;
        h.nMatch += (u32)L;
        Score += (int)L;
        Score += P.sjdbScore;
    } else {
        eA.annot = 0;
        eA.sjStr = 0;
        if (eA.iFrag == iFragB) {
            u64 gBend = gBstart + L - 1;
            u64 rBend = rBstart + L - 1;
            if (rBend <= rAend) return -1000001;
            if (gBend <= gAend) return -1000002;
            if (rBstart <= rAend) {
                gBstart += rAend - rBstart + 1;
                rBstart = rAend + 1;
                L = rBend - rBstart + 1;
            }
            Score += (int)(rBend - rBstart + 1);
            int gGap = (int)(gBstart - gAend - 1);
            int rGap = (int)(rBstart - rAend - 1);
            u64 nMatch = L, nMM = 0, Del = 0, Ins = 0, nIns = 0, nDel = 0;
            int jR = 0;
            int jCan = 999;
            u64 gBstart1 = gBstart - rGap - 1;

            if (gGap == 0 && rGap == 0) {
            } else if (gGap > 0 && rGap > 0 && rGap == gGap) {
                if constexpr (COOP) {   // matches / mismatches of the gap: two popcounts per 32 bases
                    const u32 lane = threadIdx.x & 31;
                    #pragma unroll 1
                    for (int base = 1; base <= rGap; base += 32) {
                        const int ii = base + (int)lane;
                        bool m = false, x = false;
                        if (ii <= rGap) {
                            u8 gv = Gat(ln, gAend + ii), rv = R[rAend + ii];
                            if (gv < 4 && rv < 4) { m = rv == gv; x = !m; }
                        }
                        const int nm = __popc(__ballot_sync(0xffffffffu, m)), nx = __popc(__ballot_sync(0xffffffffu, x));
                        Score += nm - nx; nMatch += (u64)nm; nMM += (u64)nx;
                    }
                } else {
                #pragma unroll 1
                for (int ii = 1; ii <= rGap; ii++) {
                    u8 gv = Gat(ln, gAend + ii), rv = R[rAend + ii];
                    if (gv < 4 && rv < 4) {
                        if (rv == gv) { Score += 1; nMatch++; } else { Score -= 1; nMM++; }
                    }
                }
                }
            } else if (gGap > rGap) {
                nDel = 1;
                Del = (u64)(gGap - rGap);
                if (Del > P.alignIntronMax && P.alignIntronMax > 0) return -1000003;
                int jPen = 0;
                u64 jjL = 0, jjR = 0;
                if constexpr (COOP) {
                    const u32 lane = threadIdx.x & 31;
                    const u32 below = (1u << lane) - 1;
                    // (1) backward scan for the start of the junction search: the sequential loop visits jR1 = 0,-1,.. and ends at the
                    // (scoreStitchSJshift+1)-th position where the read matches the donor side but not the acceptor side, or at 1-eA.L.
                    int jR1 = 0;
                    {
                        const int jMin = 1 - (int)eA.L;
                        const int need = P.scoreStitchSJshift + 1;
                        int cnt = 0;
                        #pragma unroll 1
                        for (int base = 0;; base -= 32) {
                            const int pz = base - (int)lane;
                            bool pen = false;
                            if (pz >= jMin) {
                                u8 rv = R[(i64)rAend + pz], gb = Gat(ln, gBstart1 + (u64)(i64)pz), ga = Gat(ln, gAend + (u64)(i64)pz);
                                pen = rv != gb && gb < 4 && rv == ga;
                            }
                            u32 pm = __ballot_sync(0xffffffffu, pen);
                            const int lastLane = base - jMin;          // lane holding position jMin (may be > 31)
                            int hit = -1;
                            if (need <= 0) hit = 0;
                            else if (cnt + __popc(pm) >= need) {
                                #pragma unroll 1
                                for (int q = 1; q < need - cnt; q++) pm &= pm - 1;
                                hit = __ffs(pm) - 1;
                            }
                            if (hit >= 0 && hit <= lastLane) { jR1 = base - hit; break; }
                            if (lastLane <= 31) { jR1 = jMin; break; }
                            cnt += __popc(pm);
                        }
                    }
                    // (2) forward scan: Score1 is a prefix sum of {+1,-1,0}; the junction is the FIRST position with the maximum of
                    // Score1 + motif penalty (the sequential loop updates on strict improvement only).
                    {
                        const int jEnd = int(rBend) - int(rAend);
                        const bool withMotif = Del >= P.alignIntronMin;
                        int maxScore2 = -999999;
                        int Score1 = 0;
                        #pragma unroll 1
                        for (int base = jR1; base < jEnd; base += 32) {
                            const int pz = base + (int)lane;
                            const bool v = pz < jEnd;
                            bool up = false, dn = false;
                            int jCan1 = -1, jPen1 = 0;
                            if (v) {
                                u8 rv = R[(i64)rAend + pz], ga = Gat(ln, gAend + (u64)(i64)pz), gb = Gat(ln, gBstart1 + (u64)(i64)pz);
                                up = rv == ga && rv != gb;
                                dn = rv != ga && rv == gb;
                                if (withMotif) {
                                    u8 d1 = Gat(ln, gAend + (u64)(i64)(pz + 1)), d2 = Gat(ln, gAend + (u64)(i64)(pz + 2));
                                    u8 a1 = Gat(ln, gBstart1 + (u64)(i64)(pz - 1)), a2 = gb;
                                    if (d1 == 2 && d2 == 3 && a1 == 0 && a2 == 2) { jCan1 = 1; }
                                    else if (d1 == 1 && d2 == 3 && a1 == 0 && a2 == 1) { jCan1 = 2; }
                                    else if (d1 == 2 && d2 == 1 && a1 == 0 && a2 == 2) { jCan1 = 3; jPen1 = P.scoreGapGCAG; }
                                    else if (d1 == 1 && d2 == 3 && a1 == 2 && a2 == 1) { jCan1 = 4; jPen1 = P.scoreGapGCAG; }
                                    else if (d1 == 0 && d2 == 3 && a1 == 0 && a2 == 1) { jCan1 = 5; jPen1 = P.scoreGapATAC; }
                                    else if (d1 == 2 && d2 == 3 && a1 == 0 && a2 == 3) { jCan1 = 6; jPen1 = P.scoreGapATAC; }
                                    else { jCan1 = 0; jPen1 = P.scoreGapNoncan; }
                                }
                            }
                            const u32 upM = __ballot_sync(0xffffffffu, up), dnM = __ballot_sync(0xffffffffu, dn);
                            const u32 incl = below | (1u << lane);
                            const int S2 = Score1 + __popc(upM & incl) - __popc(dnM & incl) + jPen1;
                            const int best = __reduce_max_sync(0xffffffffu, v ? S2 : (int)0x80000000);
                            if (best > maxScore2) {
                                const u32 src = (u32)__ffs(__ballot_sync(0xffffffffu, v && S2 == best)) - 1;
                                maxScore2 = best;
                                jR = base + (int)src;
                                jCan = __shfl_sync(0xffffffffu, jCan1, src);
                                jPen = __shfl_sync(0xffffffffu, jPen1, src);
                            }
                            Score1 += __popc(upM) - __popc(dnM);
                        }
                    }
                    // (3) repeat lengths around the junction: first position where the flanks differ (or an N, or 256)
                    const u64 jRu = (u64)(i64)jR;
                    #pragma unroll 1
                    for (;;) {
                        const u64 q = jjL + lane;
                        bool ok = gAend + jRu >= q && q <= 255;
                        if (ok) { u8 x = Gat(ln, gAend - q + jRu), y = Gat(ln, gBstart1 - q + jRu); ok = x == y && x < 4; }
                        const u32 fail = __ballot_sync(0xffffffffu, !ok);
                        if (fail) { jjL += (u32)__ffs(fail) - 1; break; }
                        jjL += 32;
                    }
                    #pragma unroll 1
                    for (;;) {
                        const u64 q = jjR + lane;
                        bool ok = gAend + q + jRu + 1 < g.nGenome && gBstart1 + q + jRu + 1 < g.nGenome && q <= 255;   // (beyond the genome the padding (5) ends the loop)
                        if (ok) { u8 x = Gat(ln, gAend + q + jRu + 1), y = Gat(ln, gBstart1 + q + jRu + 1); ok = x == y && x < 4; }
                        const u32 fail = __ballot_sync(0xffffffffu, !ok);
                        if (fail) { jjR += (u32)__ffs(fail) - 1; break; }
                        jjR += 32;
                    }
                } else {
                int Score1 = 0;
                int jR1 = 1;
                #pragma unroll 1
                do {
                    jR1--;
                    u8 rv = R[(i64)rAend + jR1], gb = Gat(ln, gBstart1 + (u64)(i64)jR1), ga = Gat(ln, gAend + (u64)(i64)jR1);
                    if (rv != gb && gb < 4 && rv == ga) Score1 -= 1;
                } while (Score1 + P.scoreStitchSJshift >= 0 && int(eA.L) + jR1 > 1);

                int maxScore2 = -999999;
                Score1 = 0;
                #pragma unroll 1
                do {
                    u8 rv = R[(i64)rAend + jR1], ga = Gat(ln, gAend + (u64)(i64)jR1), gb = Gat(ln, gBstart1 + (u64)(i64)jR1);
                    if (rv == ga && rv != gb) Score1 += 1;
                    if (rv != ga && rv == gb) Score1 -= 1;
                    int jCan1 = -1, jPen1 = 0, Score2 = Score1;
                    if (Del >= P.alignIntronMin) {
                        u8 d1 = Gat(ln, gAend + (u64)(i64)(jR1 + 1)), d2 = Gat(ln, gAend + (u64)(i64)(jR1 + 2));
                        u8 a1 = Gat(ln, gBstart1 + (u64)(i64)(jR1 - 1)), a2 = gb;
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

                u64 jRu = (u64)(i64)jR;
                #pragma unroll 1
                while (gAend + jRu >= jjL && Gat(ln, gAend - jjL + jRu) == Gat(ln, gBstart1 - jjL + jRu) && Gat(ln, gAend - jjL + jRu) < 4 && jjL <= 255) jjL++;
                #pragma unroll 1
                while (gAend + jjR + jRu + 1 < g.nGenome && Gat(ln, gAend + jjR + jRu + 1) == Gat(ln, gBstart1 + jjR + jRu + 1) && Gat(ln, gAend + jjR + jRu + 1) < 4 && jjR <= 255) jjR++;

                }
                if (jCan <= 0) {
                    jR -= (int)jjL;
                    if (int(eA.L) + jR < 1) return -1000005;
                    jjR += jjL;
                    jjL = 0;
                }
                {
                    int i0 = 1 < jR + 1 ? 1 : jR + 1;
                    int i1 = rGap > jR ? rGap : jR;
                    if constexpr (COOP) {
                        const u32 lane = threadIdx.x & 31;
                        #pragma unroll 1
                        for (int base = i0; base <= i1; base += 32) {
                            const int ii = base + (int)lane;
                            bool a = false, b = false, c = false;
                            if (ii <= i1) {
                                u64 g1 = (ii <= jR) ? (gAend + (u64)(i64)ii) : (gBstart1 + (u64)(i64)ii);
                                u8 gv = Gat(ln, g1), rv = R[(i64)rAend + ii];
                                if (gv < 4 && rv < 4) {
                                    const bool inGap = ii >= 1 && ii <= rGap;
                                    if (rv == gv) a = inGap; else { b = true; c = !inGap; }
                                }
                            }
                            const int na = __popc(__ballot_sync(0xffffffffu, a)), nb = __popc(__ballot_sync(0xffffffffu, b)), nc = __popc(__ballot_sync(0xffffffffu, c));
                            Score += na - nb - nc; nMatch += (u64)(i64)(na - nc); nMM += (u64)nb;
                        }
                    } else {
                    #pragma unroll 1
                    for (int ii = i0; ii <= i1; ii++) {
                        u64 g1 = (ii <= jR) ? (gAend + (u64)(i64)ii) : (gBstart1 + (u64)(i64)ii);
                        u8 gv = Gat(ln, g1), rv = R[(i64)rAend + ii];
                        if (gv < 4 && rv < 4) {
                            if (rv == gv) {
                                if (ii >= 1 && ii <= rGap) { Score += 1; nMatch++; }
                            } else {
                                Score -= 1; nMM++;
                                if (ii < 1 || ii > rGap) { Score -= 1; nMatch--; }
                            }
                        }
                    }
                    }
                }
                bool annotated = false;
                if (g.sjdbN > 0) {
                    u64 jS = gAend + (u64)(i64)jR + 1, jE = gBstart1 + (u64)(i64)jR;
                    int sjdbInd = binarySearch2(jS, jE, g.sjdbStart, g.sjdbEnd, (int)g.sjdbN);
                    if (sjdbInd >= 0) {
                        annotated = true;
                        jCan = g.sjdbMotif[sjdbInd];
                        if (g.sjdbMotif[sjdbInd] == 0) {
                            if (L <= g.sjdbShiftLeft[sjdbInd] || eA.L <= g.sjdbShiftLeft[sjdbInd]) return -1000006;
                            jR += (int)g.sjdbShiftLeft[sjdbInd];
                            if (rAend + (u64)(i64)jR >= rBend) return -1000006;
                            jjL = g.sjdbShiftLeft[sjdbInd];
                            jjR = g.sjdbShiftRight[sjdbInd];
                        }
                        eA.annot = 1;
                        eA.sjStr = g.sjdbStrand[sjdbInd];
                        Score += P.sjdbScore;
                    }
                }
                if (!annotated) {
                    if (Del >= P.alignIntronMin) {
                        Score += P.scoreGap + jPen;
                    } else {
                        Score += (int)Del * P.scoreDelBase + P.scoreDelOpen;
                        jCan = -1;
                        eA.annot = 0;
                    }
                }
                eA.shL = (u16)jjL; eA.shR = (u16)jjR;
                eA.canon = (signed char)jCan;
                if (eA.annot == 0) {
                    if (jCan > 0) eA.sjStr = (u8)(2 - jCan % 2); else eA.sjStr = 0;
                }
            } else if (__builtin_expect(rGap > gGap, 0)) {
                Ins = (u64)(rGap - gGap);
                nIns = 1;
                if (gGap == 0) {
                    jR = 0;
                } else if (gGap < 0) {
                    jR = 0;
                    Score -= (-gGap);
                } else {
                    int Score1 = 0, maxScore1 = 0;
                    #pragma unroll 1
                    for (int jR1 = 1; jR1 <= gGap; jR1++) {
                        u8 gv = Gat(ln, gAend + jR1);
                        if (gv < 4) {
                            Score1 += (R[rAend + jR1] == gv) ? 1 : -1;
                            Score1 += (R[rAend + Ins + jR1] == gv) ? -1 : +1;
                        }
                        if (Score1 > maxScore1 || (Score1 == maxScore1 && P.alignInsertionFlushRight)) { maxScore1 = Score1; jR = jR1; }
                    }
                    #pragma unroll 1
                    for (int ii = 1; ii <= gGap; ii++) {
                        u64 r1 = rAend + ii + (ii <= jR ? 0 : Ins);
                        u8 gv = Gat(ln, gAend + ii), rv = R[r1];
                        if (gv < 4 && rv < 4) {
                            if (rv == gv) { Score += 1; nMatch++; } else { Score -= 1; nMM++; }
                        }
                    }
                }
                if (P.alignInsertionFlushRight) {
                    #pragma unroll 1
                    for (; jR < (int)rBend - (int)rAend - (int)Ins; jR++) {
                        u8 gv = Gat(ln, gAend + jR + 1);
                        if (R[rAend + jR + 1] != gv || gv == 4) break;
                    }
                    if (jR == (int)rBend - (int)rAend - (int)Ins) return -1000009;
                }
                Score += (int)Ins * P.scoreInsBase + P.scoreInsOpen;
                jCan = -2;
            }

            const bool motifOk = (jCan < 0 || (jCan < 7 && nMM <= (u64)(i64)P.alignSJstitchMismatchNmax[(jCan + 1) / 2]));
            ln.memoNMM = (u32)nMM; ln.memoMotifOk = motifOk;   // (for the stitch memo of the heavy kernel)
            if ((h.nMM + nMM) <= outFilterMismatchNmaxTotal && motifOk) {
                h.nMM += (u32)nMM;
                h.nMatch += (u32)nMatch;
                if (Del >= P.alignIntronMin) { h.nGap += (u32)nDel; h.lGap += (u32)Del; }
                else { h.nDel += (u32)nDel; h.lDel += (u32)Del; }
                if (Del == 0 && Ins == 0) {
                    eA.L = (u16)(eA.L + (rBend - rAend));
                } else if (Del > 0) {
                    eA.L = (u16)((int)eA.L + jR);
                    eB.L = (u16)((i64)(rBend - rAend) - jR);
                    eB.R = (u16)((i64)rAend + jR + 1);
                    eB.G = gBstart1 + (u64)(i64)jR + 1;
                    h.nExons++;
                } else if (Ins > 0) {
                    h.nIns += (u32)nIns;
                    h.lIns += (u32)Ins;
                    eA.L = (u16)((int)eA.L + jR);
                    eB.L = (u16)((i64)(rBend - rAend) - jR - (i64)Ins);
                    eB.R = (u16)((i64)rAend + jR + (i64)Ins + 1);
                    eB.G = gAend + 1 + (u64)(i64)jR;
                    eA.canon = -2;
                    eA.annot = 0;
                    h.nExons++;
                }
            } else {
                return -1000007;
            }
        } else if (gBstart + t->ex[0].R + (u64)(i64)P.alignEndsProtrudeNbasesMax >= t->ex[0].G || t->ex[0].G < t->ex[0].R) {
            if (P.alignMatesGapMax > 0 && gBstart > eA.G + eA.L + P.alignMatesGapMax) return -1000004;
            Score += (int)L;
            ExtRes er;
            er.extendL = 0; er.maxScore = 0; er.nMatch = 0; er.nMM = 0;
            if (extendAlign<COOP>(ln, rAend + 1, gAend + 1, 1, 1, STAR_READ_SEQ_LENGTH_MAX, h.nMatch, h.nMM, outFilterMismatchNmaxTotal,
                            P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[eA.iFrag][1], er)) {
                h.maxScore += er.maxScore; h.nMatch += er.nMatch; h.nMM += er.nMM;
                Score += er.maxScore;
                eA.L = (u16)(eA.L + er.extendL);
            }
            eB.R = (u16)rBstart; eB.G = gBstart; eB.L = (u16)L;
            h.nMatch += (u32)L;
            er.extendL = 0; er.maxScore = 0; er.nMatch = 0; er.nMM = 0;
            u64 extlen = P.alignEndsTypeExt[iFragB][1] ? (u64)STAR_READ_SEQ_LENGTH_MAX : gBstart - t->ex[0].G + t->ex[0].R;
            if (extendAlign<COOP>(ln, rBstart - 1, gBstart - 1, -1, -1, extlen, h.nMatch, h.nMM, outFilterMismatchNmaxTotal,
                            P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[iFragB][1], er)) {
                h.maxScore += er.maxScore; h.nMatch += er.nMatch; h.nMM += er.nMM;
                Score += er.maxScore;
                eB.R = (u16)(eB.R - er.extendL);
                eB.G -= er.extendL;
                eB.L = (u16)(eB.L + er.extendL);
            }
            eA.canon = -3;
            eA.annot = 0;
            h.nExons++;
        } else {
            return -1000008;
        }
    }
    if constexpr (PRIV) {
        Exon& last = h.nExons == nEx0 ? eAL : eBL;
        last.iFrag = (u8)iFragB;
        last.sjA = sjAB;
        __syncwarp();                                  // every lane has finished reading the shared transcript
        if ((threadIdx.x & 31) == 0) { t->h = hL; t->ex[nEx0 - 1] = eAL; if (hL.nExons != nEx0) t->ex[nEx0] = eBL; }
        __syncwarp();
    } else {
        t->ex[h.nExons - 1].iFrag = (u8)iFragB;
        t->ex[h.nExons - 1].sjA = sjAB;
        if constexpr (COOP) __syncwarp();
    }
    return Score;
}

// blocksOverlap.cpp:3-40
__device__ u64 blocksOverlap(const DevTr& t1, const DevTr& t2) {
    u32 i1 = 0, i2 = 0;
    u64 nOverlap = 0;
    #pragma unroll 1
    while (i1 < t1.h.nExons && i2 < t2.h.nExons) {
        u64 rs1 = t1.ex[i1].R, rs2 = t2.ex[i2].R;
        u64 re1 = rs1 + t1.ex[i1].L, re2 = rs2 + t2.ex[i2].L;
        u64 gs1 = t1.ex[i1].G, gs2 = t2.ex[i2].G;
        if (rs1 >= re2) {
            i2++;
        } else if (rs2 >= re1) {
            i1++;
        } else if (gs1 - rs1 != gs2 - rs2) {
            if (re1 >= re2) i2++;
            if (re2 >= re1) i1++;
        } else {
            nOverlap += (re1 < re2 ? re1 : re2) - (rs1 > rs2 ? rs1 : rs2);
            if (re1 >= re2) i2++;
            if (re2 >= re1) i1++;
        }
    }
    return nOverlap;
}

// warp-uniform mode (flat_dfs_warp_kernel): 32-bit words [0,nWords) copied one per lane instead of 32 times by every lane
__device__ __forceinline__ void warpCopyWords(void* dst, const void* src, u32 nWords) {
    const u32 lane = threadIdx.x & 31;
    __syncwarp();
    #pragma unroll 1
    for (u32 q = lane; q < nWords; q += 32) ((u32*)dst)[q] = ((const u32*)src)[q];
    __syncwarp();
}
static_assert(sizeof(TrHead) == 80 && sizeof(Exon) == 24, "warpCopyWords call sites assume 20-word heads and 6-word exons");

__device__ __forceinline__ void copyTr(DevTr* dst, const DevTr* src) {
    dst->h = src->h;
    #pragma unroll 1
    for (u32 i = 0; i < src->h.nExons; i++) dst->ex[i] = src->ex[i];
}

__device__ int log2Score(const DevIndex& ix, u64 gLen) {
    // value of int(ceil(log2((double)gLen)*scale-0.5)) from the host-computed step table (stitchWindowAligns.cpp:221-225)
    int v = ix.log2Val[0];
    #pragma unroll 1
    for (int k = 1; k < ix.log2N; k++) {
        if (gLen >= ix.log2Thr[k]) v = ix.log2Val[k]; else break;
    }
    return v;
}

// binarySearch2 over the novel junctions of the 2nd BySJout stage (sorted by start, then end): is (jS, jE) in the list?  Cold path.
__device__ __noinline__ bool sjNovelHas(const DevIndex& g, u64 jS, u64 jE) {
    u64 lo = 0, hi = g.sjNovelN;
    while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (g.sjNovelStart[mid] < jS) lo = mid + 1; else hi = mid; }
    #pragma unroll 1
    for (; lo < g.sjNovelN && g.sjNovelStart[lo] == jS; lo++) if (g.sjNovelEnd[lo] == jE) return true;
    return false;
}

// Leaf of stitchWindowAligns, part 1 (stitchWindowAligns.cpp:19-243): extend both ends, apply the filters, compute the final score.
// Pure function of the DFS path (reads no per-read mutable state), so the heavy path can evaluate leaves of different sub-trees
// in parallel.  Result in ln.leaf (h.maxScore, h.iFrag set).  Returns false when a filter drops the transcript.
template <bool COOP = false>
__device__ bool evalLeaf(Lane& ln, int Score, u64 tR2, u64 tG2, u32 Chr, u32 Str, u32 roStr) {
    const star_params_t& P = *ln.P;
    const DevIndex& g = *ln.ix;
    DevTr& t = *ln.leaf;
    if constexpr (COOP) warpCopyWords(&t, ln.cur, 20 + 6 * (u32)ln.cur->h.nExons);   // head and exons are contiguous
    else copyTr(&t, ln.cur);
    // COOP: the leaf copy is shared by the warp: the head is worked on privately and written back by lane 0 at the end, the two exon
    // updates of the end extensions are done by lane 0 followed by a __syncwarp
    constexpr bool PRIV = COOP && SB_PRIVATE_COPIES;
    TrHead hL;
    if constexpr (PRIV) hL = t.h;
    TrHead& h = PRIV ? hL : t.h;
    const u64 Lread = ln.Lread;
    int vOrder[2];
    if (roStr == 0) { vOrder[0] = 0; vOrder[1] = 1; } else { vOrder[0] = 1; vOrder[1] = 0; }
    #pragma unroll 1
    for (int iOrd = 0; iOrd < 2; iOrd++) {
        ExtRes er;
        er.extendL = 0; er.maxScore = 0; er.nMatch = 0; er.nMM = 0;
        if (vOrder[iOrd] == 0) {
            if (h.rStart > 0) {
                u32 imate = t.ex[0].iFrag;
                if (extendAlign<COOP>(ln, (u64)h.rStart - 1, h.gStart - 1, -1, -1, h.rStart, tR2 - h.rStart + 1, h.nMM, ln.outFilterMismatchNmaxTotal,
                                P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[imate][(int)(Str != imate)], er)) {
                    h.maxScore += er.maxScore; h.nMatch += er.nMatch; h.nMM += er.nMM;
                    Score += er.maxScore;
                    h.rStart -= er.extendL;
                    h.gStart -= er.extendL;
                    if (!COOP || (threadIdx.x & 31) == 0) {
                        t.ex[0].R = (u16)h.rStart;
                        t.ex[0].G = h.gStart;
                        t.ex[0].L = (u16)(t.ex[0].L + er.extendL);
                    }
                    if constexpr (COOP) __syncwarp();
                }
            }
        } else {
            if (tR2 < Lread) {
                u32 imate = t.ex[h.nExons - 1].iFrag;
                if (extendAlign<COOP>(ln, tR2 + 1, tG2 + 1, +1, +1, Lread - tR2 - 1, tR2 - h.rStart + 1, h.nMM, ln.outFilterMismatchNmaxTotal,
                                P.outFilterMismatchNoverLmax, P.alignEndsTypeExt[imate][(int)(imate == Str)], er)) {
                    h.maxScore += er.maxScore; h.nMatch += er.nMatch; h.nMM += er.nMM;
                    Score += er.maxScore;
                    tR2 += er.extendL; tG2 += er.extendL;
                    if (!COOP || (threadIdx.x & 31) == 0) t.ex[h.nExons - 1].L = (u16)(t.ex[h.nExons - 1].L + er.extendL);
                    if constexpr (COOP) __syncwarp();
                }
            }
        }
    }
    const u32 nEx = h.nExons;
    if (!P.alignSoftClipAtReferenceEnds &&
        ((t.ex[nEx - 1].G + Lread - t.ex[nEx - 1].R) > (g.chrStart[Chr] + g.chrLength[Chr]) || t.ex[0].G < (g.chrStart[Chr] + t.ex[0].R))) return false;
    h.rLength = 0;
    #pragma unroll 1
    for (u32 i = 0; i < nEx; i++) h.rLength += t.ex[i].L;
    h.gLength = tG2 + 1 - h.gStart;
    #pragma unroll 1
    for (u32 isj = 0; isj + 1 < nEx; isj++) {
        if (t.ex[isj].canon >= 0) {
            if (t.ex[isj].annot == 1) {
                if ((t.ex[isj].L < P.alignSJDBoverhangMin && (isj == 0 || t.ex[isj - 1].canon == -3 || (t.ex[isj - 1].annot == 0 && t.ex[isj - 1].canon >= 0)))
                    || (t.ex[isj + 1].L < P.alignSJDBoverhangMin && (isj == nEx - 2 || t.ex[isj + 1].canon == -3 || (t.ex[isj + 1].annot == 0 && t.ex[isj + 1].canon >= 0))))
                    return false;
            } else {
                if (t.ex[isj].L < P.alignSJoverhangMin + t.ex[isj].shL || t.ex[isj + 1].L < P.alignSJoverhangMin + t.ex[isj].shR) return false;
            }
        }
    }
    if (nEx > 1 && t.ex[nEx - 2].annot == 1 && t.ex[nEx - 1].L < P.alignSJDBoverhangMin) return false;
    u32 sjN = 0, im[3] = {0, 0, 0};
    #pragma unroll 1
    for (u32 iex = 0; iex + 1 < nEx; iex++) {
        if (t.ex[iex].canon >= 0) { sjN++; im[t.ex[iex].sjStr]++; }
    }
    if (im[1] > 0 && im[2] == 0) h.sjMotifStrand = 1;
    else if (im[1] == 0 && im[2] > 0) h.sjMotifStrand = 2;
    else h.sjMotifStrand = 0;
    if (im[1] > 0 && im[2] > 0 && P.outFilterIntronStrandsRemoveInconsistent) return false;
    if (sjN > 0 && h.sjMotifStrand == 0 && P.outSAMstrandFieldType == 1) return false;
    if (P.outFilterIntronMotifs == 1) {
        #pragma unroll 1
        for (u32 iex = 0; iex + 1 < nEx; iex++) if (t.ex[iex].canon == 0) return false;
    } else if (P.outFilterIntronMotifs == 2) {
        #pragma unroll 1
        for (u32 iex = 0; iex + 1 < nEx; iex++) if (t.ex[iex].canon == 0 && t.ex[iex].annot == 0) return false;
    }
    {
        u64 nsj = 0, exl = 0;
        #pragma unroll 1
        for (u32 iex = 0; iex < nEx; iex++) {
            exl += t.ex[iex].L;
            if (iex == nEx - 1 || t.ex[iex].canon == -3) {
                if (nsj > 0 && (exl < P.alignSplicedMateMapLmin || exl < (u64)(P.alignSplicedMateMapLminOverLmate * (double)ln.readLength[t.ex[iex].iFrag]))) return false;
                exl = 0; nsj = 0;
            } else if (t.ex[iex].canon >= 0) {
                nsj++;
            }
        }
    }
    if (__builtin_expect(g.sjNovelOn != 0, 0)) {   // stitchWindowAligns.cpp:169-177 (2nd stage of --outFilterType BySJout)
        #pragma unroll 1
        for (u32 iex = 0; iex + 1 < nEx; iex++)
            if (t.ex[iex].canon >= 0 && t.ex[iex].annot == 0 && !sjNovelHas(g, t.ex[iex].G + t.ex[iex].L, t.ex[iex + 1].G - 1)) return false;
    }
    if (t.ex[0].iFrag != t.ex[nEx - 1].iFrag) {
        if (t.ex[nEx - 1].G + t.ex[nEx - 1].L <= t.ex[0].G) return false;
        u32 iexM2 = nEx;
        #pragma unroll 1
        for (u32 iex = 0; iex + 1 < nEx; iex++) {
            if (t.ex[iex].canon == -3) { iexM2 = iex + 1; break; }
        }
        if (t.ex[iexM2 - 1].G + t.ex[iexM2 - 1].L > t.ex[iexM2].G) {
            if (t.ex[0].G > t.ex[iexM2].G + t.ex[0].R + (u64)(i64)P.alignEndsProtrudeNbasesMax) return false;
            if (t.ex[iexM2 - 1].G + t.ex[iexM2 - 1].L > t.ex[nEx - 1].G + Lread - t.ex[nEx - 1].R + (u64)(i64)P.alignEndsProtrudeNbasesMax) return false;
            u32 iex1 = 1, iex2 = iexM2 + 1;
            #pragma unroll 1
            for (; iex1 < iexM2; iex1++) {
                if (t.ex[iex1].G >= t.ex[iex2 - 1].G + t.ex[iex2 - 1].L) break;
            }
            #pragma unroll 1
            while (iex1 < iexM2 && iex2 < nEx) {
                if (t.ex[iex1 - 1].canon < 0) { iex1++; continue; }
                if (t.ex[iex2 - 1].canon < 0) { iex2++; continue; }
                if ((t.ex[iex1].G != t.ex[iex2].G) || ((t.ex[iex1 - 1].G + t.ex[iex1 - 1].L) != (t.ex[iex2 - 1].G + t.ex[iex2 - 1].L))) return false;
                iex1++; iex2++;
            }
        }
    }
    if (P.scoreGenomicLengthLog2scale != 0) {
        Score += log2Score(g, t.ex[nEx - 1].G + t.ex[nEx - 1].L - t.ex[0].G);
        Score = Score > 0 ? Score : 0;
    }
    h.maxScore = Score;
    h.iFrag = (t.ex[0].iFrag == t.ex[nEx - 1].iFrag) ? (signed char)t.ex[0].iFrag : (signed char)-1;
    if constexpr (PRIV) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) t.h = hL;
        __syncwarp();
    } else if constexpr (COOP) {
        __syncwarp();
    }
    return true;
}

// Leaf, part 2 (stitchWindowAligns.cpp:228-304): maxScoreMate update, record test, dedup by blocksOverlap, ordered insert.
// Strictly sequential in DFS order within a window and in window order within a read.
template <bool COOP = false>
__device__ void recordLeaf(Lane& ln, u16* wTr, u16* nWinTr) {
    const star_params_t& P = *ln.P;
    DevTr& t = *ln.leaf;
    TrHead& h = t.h;
    const u32 nEx = h.nExons;
    const int Score = h.maxScore;
    if (h.iFrag >= 0) {
        if (ln.maxScoreMate[h.iFrag] < Score) ln.maxScoreMate[h.iFrag] = Score;
    }
    int wBest = ln.pool[wTr[0]].h.maxScore;
    if (Score + P.outFilterMultimapScoreRange >= wBest || (h.iFrag >= 0 && Score + P.outFilterMultimapScoreRange >= ln.maxScoreMate[h.iFrag])) {
        u32 iTr = 0;
        // COOP (warp-uniform caller): the leaf, the slot table and the pool are shared by the warp, so lane 0 is the only writer and
        // every write is followed by a __syncwarp before any lane reads it
        u32 mappedLength = 0;
        #pragma unroll 1
        for (u32 iex = 0; iex < nEx; iex++) mappedLength += t.ex[iex].L;
        if constexpr (COOP) {
            if ((threadIdx.x & 31) == 0) h.mappedLength = mappedLength;
            __syncwarp();
        } else {
            h.mappedLength = mappedLength;
        }
        u32 n = *nWinTr;
        #pragma unroll 1
        while (iTr < n) {
            const DevTr& o = ln.pool[wTr[iTr]];
            u64 nOverlap = blocksOverlap(t, o);
            u64 uNew = mappedLength - nOverlap;
            u64 uOld = o.h.mappedLength - nOverlap;
            if (uNew == 0 && Score < o.h.maxScore) {
                break;
            } else if (uOld == 0) {
                if constexpr (COOP) __syncwarp();
                if (!COOP || (threadIdx.x & 31) == 0) {
                    u16 p = wTr[iTr];
                    #pragma unroll 1
                    for (u32 ii = iTr + 1; ii < n; ii++) wTr[ii - 1] = wTr[ii];
                    wTr[n - 1] = p;
                }
                if constexpr (COOP) __syncwarp();
                n--;
            } else if (uOld > 0 && (uNew > 0 || Score >= o.h.maxScore)) {
                iTr++;
            }
        }
        if (iTr == n) {
            #pragma unroll 1
            for (iTr = 0; iTr < n; iTr++) {
                const DevTr& o = ln.pool[wTr[iTr]];
                if (Score > o.h.maxScore || (Score == o.h.maxScore && h.gLength < o.h.gLength)) break;
            }
            u16 p = wTr[n];
            if constexpr (COOP) __syncwarp();          // (every lane has read wTr[n] and finished the scans above)
            if (!COOP || (threadIdx.x & 31) == 0) {
                #pragma unroll 1
                for (int ii = (int)n; ii > (int)iTr; ii--) wTr[ii] = wTr[ii - 1];
                wTr[iTr] = p;
            }
            if constexpr (COOP) warpCopyWords(&ln.pool[p], &t, 20 + 6 * nEx);   // (one word per lane; syncs before and after)
            else copyTr(&ln.pool[p], &t);
            if (n < P.alignTranscriptsPerWindowNmax) n++;
        }
        *nWinTr = (u16)n;
    }
}

// ---- stitch memo (heavy kernel).  Inside one window the include/exclude enumeration stitches the same ordered pair of seeds
// (A = last included seed, B = candidate) over and over under different earlier choices.  For seeds of the same mate the result of
// stitchAlignToTranscript is a pure function of (A, B, current length of A's exon) except for the final test on the TOTAL number
// of mismatches, which is re-evaluated on every use.  Entries are published with a seqlock-style key so that lanes of the warp
// can share the table without locks; a lost race only costs a recomputation.
struct StitchMemo {
    volatile u64 key;     // 0 = empty / being written
    int dScore;           // <= -1000000: the pair never stitches (independent of the path)
    u32 nMM, nMatch;
    u32 dGapN, dGapL, dDelN, dDelL, dInsN, dInsL;
    u64 eB_G;
    u16 eB_R, eB_L, eA_L, shL, shR;
    signed char canon;
    u8 annot, sjStr;
};

static_assert(sizeof(StitchMemo) == 72, "engine_api.cu sizes the memo table with 72-byte entries");

template <bool COOP = false>
__device__ int stitchMemoized(Lane& ln, u64 rAend, u64 gAend, const Seed& s, u32 bIdx, DevTr* t, bool& wasHit) {
    wasHit = false;
    TrHead& h = t->h;
    if (!ln.memo || ln.lastSeed < 0 || h.nExons >= STAR_MAX_N_EXONS || t->ex[h.nExons - 1].iFrag != s.iFrag)
        return stitchAlignToTranscript<COOP>(ln, rAend, gAend, s.rStart, s.gStart, s.Length, s.iFrag, s.sjA, t);
    Exon& eA = t->ex[h.nExons - 1];
    {   // the two cheapest outcomes are not worth a table access: B ends inside A in read or genome space (:53-54).  The sjdb shortcut
        // (:18) is tested first by the reference, so it must not apply here.
        const bool sjdbDirect = s.sjA != SJA_NONE && eA.sjA == s.sjA && (u64)s.rStart == rAend + 1 && gAend + 1 < s.gStart;
        if (!sjdbDirect) {
            if ((u64)s.rStart + s.Length - 1 <= rAend) { eA.annot = 0; eA.sjStr = 0; return -1000001; }
            if (s.gStart + s.Length - 1 <= gAend) { eA.annot = 0; eA.sjStr = 0; return -1000002; }
        }
    }
    const u64 key = ln.memoBase | ((u64)(u32)ln.lastSeed << 22) | ((u64)bIdx << 16) | (u64)eA.L;
    StitchMemo* m = ln.memo + (u32)((key * 0x9E3779B97F4A7C15ULL) >> 40 & ln.memoMask);
    if (m->key == key) {
        const int dScore = m->dScore;
        const u32 nMM = m->nMM, nMatch = m->nMatch, dGapN = m->dGapN, dGapL = m->dGapL, dDelN = m->dDelN, dDelL = m->dDelL, dInsN = m->dInsN, dInsL = m->dInsL;
        const u64 eB_G = m->eB_G;
        const u16 eB_R = m->eB_R, eB_L = m->eB_L, eA_L = m->eA_L, shL = m->shL, shR = m->shR;
        const signed char canon = m->canon; const u8 annot = m->annot, sjStr = m->sjStr;
        __threadfence_block();
        if (m->key == key) {   // the entry was not replaced while it was read
            wasHit = true;
            ln.memoHit++;
            if (dScore <= -1000000) return dScore;
            if (h.nMM + nMM > ln.outFilterMismatchNmaxTotal) return -1000007;   // the only path-dependent test (stitchAlignToTranscript.cpp:314)
            Exon& eB = t->ex[h.nExons];
            eA.L = eA_L; eA.canon = canon; eA.annot = annot; eA.sjStr = sjStr; eA.shL = shL; eA.shR = shR;
            eB.R = eB_R; eB.G = eB_G; eB.L = eB_L; eB.iFrag = s.iFrag; eB.sjA = s.sjA;
            h.nMM += nMM; h.nMatch += nMatch; h.nGap += dGapN; h.lGap += dGapL; h.nDel += dDelN; h.lDel += dDelL; h.nIns += dInsN; h.lIns += dInsL;
            h.nExons++;
            return dScore;
        }
    }
    // miss: compute, then publish
    ln.memoMiss++;
    const TrHead h0 = h;
    ln.memoMotifOk = true; ln.memoNMM = 0;
    const int dScore = stitchAlignToTranscript<COOP>(ln, rAend, gAend, s.rStart, s.gStart, s.Length, s.iFrag, s.sjA, t);
    bool cache = true;
    if (dScore == -1000007 && ln.memoMotifOk) cache = false;   // failed only because of the total mismatch count of THIS path
    if (cache && atomicExch((unsigned long long*)&m->key, 1ULL) != 1ULL) {   // 1 = slot locked by another lane: skip publishing
        __threadfence_block();
        m->dScore = dScore;
        if (dScore > -1000000) {
            const Exon& a = t->ex[h0.nExons - 1];
            const Exon& b = t->ex[h0.nExons];
            m->nMM = h.nMM - h0.nMM; m->nMatch = h.nMatch - h0.nMatch;
            m->dGapN = h.nGap - h0.nGap; m->dGapL = h.lGap - h0.lGap; m->dDelN = h.nDel - h0.nDel; m->dDelL = h.lDel - h0.lDel;
            m->dInsN = h.nIns - h0.nIns; m->dInsL = h.lIns - h0.lIns;
            m->eB_G = b.G; m->eB_R = b.R; m->eB_L = b.L; m->eA_L = a.L; m->shL = a.shL; m->shR = a.shR;
            m->canon = a.canon; m->annot = a.annot; m->sjStr = a.sjStr;
        }
        __threadfence_block();
        m->key = key;
    }
    return dScore;
}

// stitchWindowAligns.cpp:8-353 as an explicit DFS with undo records (see file header).
// dfsInit starts a window; dfsStep runs the cheap bookkeeping transitions (undo, pop) until it has executed ONE seed-include
// attempt (one stitchAlignToTranscript call), reached a leaf, or emptied the stack.  One call = one unit of work of the
// warp-lockstep state machine in stitch_kernel: all lanes of a warp that are inside a window execute their stitch call together.
template <bool COOP = false>
__device__ __forceinline__ void dfsInit(Lane& ln) {
    // trA = *trInit with Chr/Str set (ReadAlign_stitchPieces.cpp:282-286)
    if constexpr (COOP) {
        __syncwarp();
        if ((threadIdx.x & 31) < 20) ((u32*)&ln.cur->h)[threadIdx.x & 31] = 0;   // every field of the initial head is zero
        __syncwarp();
        ln.inclMask = 0;
        ln.level = 0; ln.nInc = 0; ln.Score = 0; ln.tR2 = 0; ln.tG2 = 0;
        ln.lastSeed = -1;
        ln.ph[0] = 0;
        return;
    }
    TrHead z;
    z.gStart = 0; z.gLength = 0; z.rStart = 0; z.rLength = 0; z.maxScore = 0; z.nMatch = 0; z.nMM = 0; z.mappedLength = 0;
    z.nGap = 0; z.lGap = 0; z.nDel = 0; z.lDel = 0; z.nIns = 0; z.lIns = 0; z.nUnique = 0; z.nAnchor = 0; z.nExons = 0; z.iFrag = 0;
    z.sjMotifStrand = 0; z.primaryFlag = 0; z.pad[0] = z.pad[1] = z.pad[2] = 0;
    ln.cur->h = z;
    ln.inclMask = 0;
    ln.level = 0; ln.nInc = 0; ln.Score = 0; ln.tR2 = 0; ln.tG2 = 0;
    ln.lastSeed = -1;
    ln.ph[0] = 0;
}

// after a sub-tree is exhausted: unwind to the deepest level whose exclude branch is still unexplored (undoing its include)
template <bool COOP = false>
__device__ __forceinline__ void dfsBacktrack(Lane& ln) {
    DevTr* t = ln.cur;
    // The deepest level whose exclude branch is still open is the last included seed of the path (levels above it were excluded on
    // the way down), unless that include was forced by the task prefix (ph 3): then the sub-tree is exhausted.
    const int L = ln.lastSeed;
    if (L >= 0 && ln.ph[L] == 1) {
        // undo the include of seed L, then explore the branch without it (WA_Anchor==2 never occurs: WlastAnchor is initialised
        // to (uint)-1 and only updated when WlastAnchor<iA, ReadAlign_stitchPieces.cpp:117, assignAlignToWindow.cpp:128)
        const Frame& u = ln.stack[--ln.nInc];
        if constexpr (COOP) {
            if (u.h.nExons > 0) warpCopyWords(&t->ex[u.h.nExons - 1], &u.last, 6);
            warpCopyWords(&t->h, &u.h, 20);
        } else {
            if (u.h.nExons > 0) t->ex[u.h.nExons - 1] = u.last;
            t->h = u.h;
        }
        ln.Score = u.Score; ln.tR2 = u.tR2; ln.tG2 = u.tG2;
        ln.lastSeed = (int)u.pad[0] - 1;
        ln.inclMask &= ~(1ULL << L);
        ln.ph[L] = 2;
        ln.level = L + 1;
        ln.ph[L + 1] = 0;
        return;
    }
    ln.level = -1;
}

#define DFS_CONTINUE 0
#define DFS_LEAF 1
#define DFS_DONE 2
template <bool COOP = false>
__device__ int dfsStep(Lane& ln, const Seed* __restrict__ WA, u32 nA) {
    DevTr* t = ln.cur;
    int runAhead = 0;
    #pragma unroll 1
    for (;;) {
        if (ln.level < 0) return DFS_DONE;
        const u32 L = (u32)ln.level;
        ln.nodes++;
        if (L >= nA) {
            if (ln.ph[L] == 9) {               // the leaf at this position was already handed out: now unwind
                dfsBacktrack<COOP>(ln);
                continue;
            }
            const bool isLeaf = ln.tR2 != 0;   // "iA>=nA && tR2==0: no aligns in the transcript" (:14)
            if (isLeaf) {
                // the caller evaluates the leaf from ln.cur, so the undo of the last include has to wait for the next call
                ln.leafScore = ln.Score; ln.leafR2 = ln.tR2; ln.leafG2 = ln.tG2;
                ln.ph[L] = 9;
                ln.nodes--;                    // (this level is visited twice)
                return DFS_LEAF;
            }
            dfsBacktrack<COOP>(ln);
            continue;
        }
        const bool forced = L < ln.forceDepth;
        if (forced && ((ln.forceBits >> (ln.forceDepth - 1 - L)) & 1u)) {   // this level is fixed to "exclude"
            ln.ph[L] = 2;
            ln.level = (int)L + 1;
            ln.ph[L + 1] = 0;
            continue;
        }
        if constexpr (COOP) {
            // warp-uniform mode: the quick-fail test below for seeds L, L+1, .. is independent of the outcome for the earlier ones (nothing
            // is included in between), so 32 seeds are tested at once and the DFS jumps to the first seed that needs a real attempt.
            if (t->h.nExons > 0 && !forced) {
                const u32 lane = threadIdx.x & 31;
                const Exon& eA = t->ex[t->h.nExons - 1];
                const bool full = t->h.nExons >= STAR_MAX_N_EXONS;
                const u32 eFrag = eA.iFrag, eSj = eA.sjA;
                u32 j = L;
                #pragma unroll 1
                for (;;) {
                    const u32 idx = j + lane;
                    bool real = false;
                    if (idx < nA) {
                        const Seed q = WA[idx];
                        bool qf = full;
                        if (!qf && eFrag == q.iFrag) {
                            const bool sjdbDirect = q.sjA != SJA_NONE && eSj == q.sjA && (u64)q.rStart == (u64)ln.tR2 + 1 && ln.tG2 + 1 < q.gStart;
                            qf = !sjdbDirect && ((u64)q.rStart + q.Length - 1 <= ln.tR2 || q.gStart + q.Length - 1 <= ln.tG2);
                        }
                        real = !qf;
                    }
                    const u32 realMask = __ballot_sync(0xffffffffu, real);
                    const u32 span = nA - j < 32 ? nA - j : 32;
                    const u32 skipped = realMask ? (u32)__ffs(realMask) - 1 : span;
                    if (lane < skipped) ln.ph[j + lane] = 2;
                    j += skipped;
                    if (realMask || j >= nA) break;
                }
                __syncwarp();
                if (j > L) {
                    ln.nodes += (u64)(j - L - 1);   // (the skipped levels are visited one by one in the sequential order; this one is counted above)
                    ln.level = (int)j;
                    ln.ph[j] = 0;
                    continue;
                }
            }
        }
        const Seed s = WA[L];
        if (t->h.nExons > 0) {
            // The most frequent outcomes of an include attempt change nothing: the transcript is full (:13) or seed B ends inside the
            // last included seed in read or genome space (:53-54; tested after the sjdb shortcut :18).  Decide them before any state is saved.
            const Exon& eA = t->ex[t->h.nExons - 1];
            bool quickFail = t->h.nExons >= STAR_MAX_N_EXONS;
            if (!quickFail && eA.iFrag == s.iFrag) {
                const bool sjdbDirect = s.sjA != SJA_NONE && eA.sjA == s.sjA && (u64)s.rStart == (u64)ln.tR2 + 1 && ln.tG2 + 1 < s.gStart;
                quickFail = !sjdbDirect && ((u64)s.rStart + s.Length - 1 <= ln.tR2 || s.gStart + s.Length - 1 <= ln.tG2);
            }
            if (quickFail) {
                if (forced) { ln.level = -1; return DFS_DONE; }
                ln.ph[L] = 2;
                ln.level = (int)L + 1;
                ln.ph[L + 1] = 0;
                if (++runAhead < 64) continue;
                return DFS_CONTINUE;
            }
        }
        Frame& u = ln.stack[ln.nInc];
        if constexpr (COOP) {
            warpCopyWords(&u.h, &t->h, 20);
            if (t->h.nExons > 0) warpCopyWords(&u.last, &t->ex[t->h.nExons - 1], 6);
        } else {
            u.h = t->h;
            if (t->h.nExons > 0) u.last = t->ex[t->h.nExons - 1];
        }
        u.Score = ln.Score; u.tR2 = ln.tR2; u.tG2 = ln.tG2;
        u.pad[0] = (u32)(ln.lastSeed + 1);
        int dScore = 0;
        bool cheap = false;
        if (t->h.nExons > 0) {
            if constexpr (COOP) dScore = stitchAlignToTranscript<true>(ln, ln.tR2, ln.tG2, s.rStart, s.gStart, s.Length, s.iFrag, s.sjA, t);   // (no stitch memo on this path)
            else dScore = stitchMemoized<COOP>(ln, ln.tR2, ln.tG2, s, L, t, cheap);
        } else {
            t->ex[0].R = s.rStart; t->h.rStart = s.rStart;
            t->ex[0].G = s.gStart; t->h.gStart = s.gStart;
            t->ex[0].L = s.Length; t->ex[0].iFrag = s.iFrag; t->ex[0].sjA = s.sjA;
            t->ex[0].canon = 0; t->ex[0].annot = 0; t->ex[0].sjStr = 0; t->ex[0].shL = 0; t->ex[0].shR = 0;
            t->h.nExons = 1;
            dScore = s.Length;
            t->h.nMatch = s.Length;
            cheap = true;
        }
        if (dScore > -1000000) {
            if (s.Nrep == 1) t->h.nUnique++;
            if (s.Anchor > 0) t->h.nAnchor++;
            ln.inclMask |= 1ULL << L;
            ln.nInc++;
            ln.lastSeed = (int)L;
            ln.ph[L] = forced ? 3 : 1;   // a forced include never explores its exclude branch
            ln.Score += dScore; ln.tR2 = (u32)s.rStart + s.Length - 1; ln.tG2 = s.gStart + s.Length - 1;
        } else {
            if (forced) { ln.level = -1; return DFS_DONE; }   // the fixed prefix is not a valid path: this sub-tree is empty
            if constexpr (COOP) {
                if (u.h.nExons > 0) warpCopyWords(&t->ex[u.h.nExons - 1], &u.last, 6);
                warpCopyWords(&t->h, &u.h, 20);
            } else {
            if (u.h.nExons > 0) t->ex[u.h.nExons - 1] = u.last;   // the failed attempt may have touched the last exon / head
            t->h = u.h;
            }
            ln.ph[L] = 2;
        }
        ln.level = (int)L + 1;
        ln.ph[L + 1] = 0;
        if (cheap && ++runAhead < 64) continue;   // memo hits / first seeds cost nothing: keep going inside this step
        return DFS_CONTINUE;
    }
}

// sjAlignSplit.cpp:3-15
__device__ __forceinline__ bool sjAlignSplit(const DevIndex& g, u64 a1, u64 aLength, u64& a1D, u64& aLengthD, u64& a1A, u64& aLengthA, u32& isj) {
    u64 sj1 = (a1 - g.sjGstart) % g.sjdbLength;
    if (sj1 < g.sjdbOverhang && sj1 + aLength > g.sjdbOverhang) {
        u64 j = (a1 - g.sjGstart) / g.sjdbLength;
        isj = (u32)j;
        aLengthD = g.sjdbOverhang - sj1;
        aLengthA = aLength - aLengthD;
        a1D = g.sjDstart[j] + sj1;
        a1A = g.sjAstart[j];
        return true;
    }
    return false;
}

__device__ __forceinline__ u32 chrOfBin(const Lane& ln, u64 bin) {
    u64 cb = bin >> ln.P->winBinChrNbits;
    return cb < ln.ix->chrBinN ? ln.ix->chrBin[cb] : 0xFFFFFFFFu;
}

// ReadAlign_createExtendWindowsWithAlign.cpp:7-84 on the interval list (see file header)
__device__ int createExtendWindowsWithAlign(Lane& ln, u64 a1, u32 aStr) {
    const star_params_t& P = *ln.P;
    u64 aBin = a1 >> P.winBinNbits;
    Window* W = ln.win;
    const u32 nW = ln.nW;
    int left = -1, right = -1;
    u64 lo = aBin > P.winAnchorDistNbins ? aBin - P.winAnchorDistNbins : 0;
    u64 hiX = aBin + P.winAnchorDistNbins + 1 < P.winBinN ? aBin + P.winAnchorDistNbins + 1 : P.winBinN;   // exclusive
    u64 bestL = 0, bestR = 0;
    #pragma unroll 1
    for (u32 w = 0; w < nW; w++) {
        if (W[w].Str != aStr || W[w].gStart > W[w].gEnd) continue;
        u64 s = W[w].gStart, e = W[w].gEnd;
        if (s <= aBin && aBin <= e) return 0;   // the bin already belongs to a window
        if (aBin > 0 && e < aBin && e >= lo) { if (left < 0 || e > bestL) { left = (int)w; bestL = e; } }
        if (aBin + 1 < P.winBinN && s > aBin && s < hiX) { if (right < 0 || s < bestR) { right = (int)w; bestR = s; } }
    }
    bool flagMergeLeft = left >= 0 && chrOfBin(ln, bestL) == chrOfBin(ln, aBin);
    bool flagMergeRight = right >= 0 && chrOfBin(ln, bestR) == chrOfBin(ln, aBin);
    u64 iBinLeft = aBin, iBinRight = aBin;
    int iWin = -1;
    if (flagMergeLeft) { iWin = left; iBinLeft = W[left].gStart; }
    if (flagMergeRight) { iBinRight = W[right].gEnd; if (!flagMergeLeft) iWin = right; }
    if (!flagMergeLeft && !flagMergeRight) {
        if (nW >= ln.caps.maxW) { ln.overflow = 1; return 101; }   // reason 1: windows
        Window nw;
        nw.gStart = (u32)aBin; nw.gEnd = (u32)aBin; nw.Chr = chrOfBin(ln, aBin); nw.nWA = 0; nw.WALrec = 0; nw.Str = (u8)aStr;
        nw.pad[0] = nw.pad[1] = nw.pad[2] = 0;
        W[nW] = nw;
        ln.nW = nW + 1;
        if (ln.nW >= P.alignWindowsPerReadNmax) {
            ln.nW = (u32)P.alignWindowsPerReadNmax - 1;
            return 101;   // EXIT_createExtendWindowsWithAlign_TOO_MANY_WINDOWS
        }
    } else {
        W[iWin].gStart = (u32)iBinLeft;
        W[iWin].gEnd = (u32)iBinRight;
        if (flagMergeLeft && flagMergeRight) { W[right].gStart = 1; W[right].gEnd = 0; }
    }
    return 0;
}

__device__ __forceinline__ void setSeed(Seed& d, u64 a1, u64 aLength, u64 aNrep, u32 aFrag, u64 aRstart, bool aAnchor, u32 sjA) {
    d.gStart = a1; d.sjA = sjA; d.rStart = (u16)aRstart; d.Length = (u16)aLength; d.Nrep = (u16)aNrep; d.Anchor = aAnchor ? 1 : 0; d.iFrag = (u8)aFrag;
}

// ReadAlign_assignAlignToWindow.cpp:6-130.  Returns false when the read hit MARKER_TOO_MANY_ANCHORS_PER_WINDOW.
__device__ bool assignAlignToWindow(Lane& ln, u64 a1, u64 aLength, u32 aStr, u64 aNrep, u32 aFrag, u64 aRstart, bool aAnchor, u32 sjA) {
    const star_params_t& P = *ln.P;
    u64 bin = a1 >> P.winBinNbits;
    int iW = -1;
    #pragma unroll 1
    for (u32 w = 0; w < ln.nW; w++) {
        if (ln.win[w].Str == aStr && ln.win[w].gStart <= bin && bin <= ln.win[w].gEnd) { iW = (int)w; break; }
    }
    if (iW < 0) return true;
    Window& W = ln.win[iW];
    if (!aAnchor && aLength < W.WALrec) return true;
    Seed* WA = ln.wa + (u64)iW * ln.caps.spw;
    u32 nWA = W.nWA;
    {
        u32 iA;
        #pragma unroll 1
        for (iA = 0; iA < nWA; iA++) {
            const Seed& s = WA[iA];
            if (aFrag == s.iFrag && s.sjA == sjA && a1 + s.rStart == s.gStart + aRstart
                && ((aRstart >= s.rStart && aRstart < (u64)s.rStart + s.Length) || (aRstart + aLength >= s.rStart && aRstart + aLength < (u64)s.rStart + s.Length))) break;
        }
        if (iA < nWA) {
            if (aLength > WA[iA].Length) {
                u32 iA0;
                #pragma unroll 1
                for (iA0 = 0; iA0 < nWA; iA0++) {
                    if (iA0 != iA && aRstart < WA[iA0].rStart) break;
                }
                if (iA0 > iA) --iA0;
                if (iA0 < iA) {
                    #pragma unroll 1
                    for (u32 iA1 = iA; iA1 > iA0; iA1--) WA[iA1] = WA[iA1 - 1];
                } else if (iA0 > iA) {
                    #pragma unroll 1
                    for (u32 iA1 = iA; iA1 < iA0; iA1++) WA[iA1] = WA[iA1 + 1];
                }
                setSeed(WA[iA0], a1, aLength, aNrep, aFrag, aRstart, aAnchor, sjA);
            }
            return true;
        }
    }
    if (nWA == P.seedPerWindowNmax) {
        u32 rec = ln.Lread + 1;
        #pragma unroll 1
        for (u32 iA = 0; iA < nWA; iA++) if (WA[iA].Anchor != 1 && WA[iA].Length < rec) rec = WA[iA].Length;
        W.WALrec = (u16)rec;
        if (rec == ln.Lread + 1) return false;   // mapMarker=MARKER_TOO_MANY_ANCHORS_PER_WINDOW; nW=0
        if (!aAnchor && aLength < rec) return true;
        u32 iA1 = 0;
        #pragma unroll 1
        for (u32 iA = 0; iA < nWA; iA++) {
            if (WA[iA].Anchor == 1 || WA[iA].Length > rec) { WA[iA1] = WA[iA]; iA1++; }
        }
        nWA = iA1;
        W.nWA = (u16)nWA;
    }
    if (aAnchor || aLength > W.WALrec) {
        u32 iA;
        #pragma unroll 1
        for (iA = 0; iA < nWA; iA++) if (aRstart < WA[iA].rStart) break;
        #pragma unroll 1
        for (u32 iA1 = nWA; iA1 > iA; iA1--) WA[iA1] = WA[iA1 - 1];
        setSeed(WA[iA], a1, aLength, aNrep, aFrag, aRstart, aAnchor, sjA);
        W.nWA = (u16)(nWA + 1);
    }
    return true;
}

__device__ void exportAlign(const DevTr& t, u32 Chr, u32 Str, u32 roStr, u32 Lread, const DevIndex& g, star_align_t* o) {
    const u32 n = t.h.nExons;
    #pragma unroll 1
    for (u32 i = 0; i < STAR_MAX_N_EXONS; i++) {
        bool v = i < n;
        bool j = i + 1 < n;
        o->exG[i] = v ? t.ex[i].G : 0; o->exR[i] = v ? t.ex[i].R : 0; o->exL[i] = v ? t.ex[i].L : 0; o->exFrag[i] = v ? t.ex[i].iFrag : 0;
        o->canonSJ[i] = j ? t.ex[i].canon : 0; o->sjAnnot[i] = j ? t.ex[i].annot : 0; o->sjStr[i] = j ? t.ex[i].sjStr : 0;
        bool sh = j && t.ex[i].canon >= -1;
        o->shiftSJ[i][0] = sh ? t.ex[i].shL : 0; o->shiftSJ[i][1] = sh ? t.ex[i].shR : 0;
    }
    o->nExons = n; o->Chr = Chr; o->Str = (u8)Str; o->roStr = (u8)roStr; o->primaryFlag = t.h.primaryFlag; o->sjMotifStrand = t.h.sjMotifStrand;
    o->iFrag = t.h.iFrag; o->maxScore = t.h.maxScore; o->nMatch = t.h.nMatch; o->nMM = t.h.nMM; o->nGap = t.h.nGap; o->lGap = t.h.lGap;
    o->nDel = t.h.nDel; o->lDel = t.h.lDel; o->nIns = t.h.nIns; o->lIns = t.h.lIns; o->nUnique = t.h.nUnique; o->nAnchor = t.h.nAnchor;
    o->rStart = t.h.rStart; o->rLength = t.h.rLength;
    o->roStart = roStr == 0 ? t.h.rStart : Lread - t.h.rStart - t.h.rLength;   // ReadAlign_multMapSelect.cpp:52
    o->gStart = t.h.gStart; o->gLength = t.h.gLength; o->cStart = t.h.gStart - g.chrStart[Chr];
}

// ------------------------------------------------------------------------------------------------------------------
// Per-read stitching bookkeeping shared by the light kernel (one lane per read) and the heavy kernel (one warp per read).

__device__ __forceinline__ void readBegin(Lane& ln, const ReadInfo& ri) {
    ln.saEnum = 0; ln.nodes = 0; ln.leaves = 0; ln.overflow = 0;
    ln.Lread = ri.Lread; ln.readLength[0] = ri.readLength[0]; ln.readLength[1] = ri.readLength[1];
    ln.outFilterMismatchNmaxTotal = ri.outFilterMismatchNmaxTotal;
    ln.maxScoreMate[0] = 0; ln.maxScoreMate[1] = 0;
    ln.trNtotal = 0; ln.nW1 = 0; ln.bestPool = -1; ln.bestScore = 0; ln.bestGLength = 0;
    ln.forceDepth = 0; ln.forceBits = 0;
}

// start of a window's stitching (ReadAlign_stitchPieces.cpp:281-294).  Returns 0 ok, 1 reference's per-read transcript budget reached
// (:288-292, remaining windows are skipped), 2 this lane's pool is full (overflow tier).
__device__ __forceinline__ int windowBegin(Lane& ln, u16*& wTr, u16& nWinTr, bool writer = true) {   // writer: false on the lanes of a warp-uniform caller that do not own the shared state
    const star_params_t& P = *ln.P;
    if (ln.trNtotal + P.alignTranscriptsPerWindowNmax >= P.alignTranscriptsPerReadNmax) return 1;
    if (ln.trNtotal + 1 > ln.caps.maxTr) return 2;
    wTr = ln.trPtr + ln.trNtotal;
    nWinTr = 0;
    // *(trAll[iW1][0]) = trA : the window-best comparison starts from maxScore 0 (:293)
    if (writer) {
        ln.pool[wTr[0]].h.maxScore = 0;
        ln.pool[wTr[0]].h.nExons = 0;
    }
    return 0;
}

// end of a window (:324-331)
__device__ __forceinline__ void windowEnd(Lane& ln, u32 Chr, u32 Str, const u16* wTr, u16 nWinTr, bool writer = true) {
    if (nWinTr == 0) return;
    const TrHead& b = ln.pool[wTr[0]].h;
    if (b.maxScore > ln.bestScore || (b.maxScore == ln.bestScore && b.gLength < ln.bestGLength)) {
        ln.bestPool = wTr[0]; ln.bestScore = b.maxScore; ln.bestGLength = b.gLength;
    }
    if (writer) {
        ln.winBase[ln.nW1] = (u16)ln.trNtotal;
        ln.winN[ln.nW1] = nWinTr;
        ln.win[ln.nW1].Chr = Chr;      // compact (iW1 <= iW): Chr/Str of the windows that have transcripts
        ln.win[ln.nW1].Str = (u8)Str;
    }
    ln.trNtotal += nWinTr;
    ln.nW1++;
}

// multMapSelect (ReadAlign_multMapSelect.cpp:8-95) + mappedFilter (ReadAlign_mappedFilter.cpp:3-20) + export of the selected alignments
__device__ void selectExport(Lane& ln, ReadInfo& ri, u32 i, u32 mapMarker, u32 bestRLength, star_read_result_t* __restrict__ results,
                             star_align_t* __restrict__ staged, ReadInfo* __restrict__ info) {
    const star_params_t& P = *ln.P;
    const DevIndex& ix = *ln.ix;
    const u32 Lread = ri.Lread;
    star_read_result_t res;
    res.unmapType = 0; res.nTr = 0; res.nTrOut = 0; res.mapMarker = 0; res.trOffset = 0; res.bestScore = 0; res.bestNMM = 0;
    res.bestRLength = 0; res.Lread = Lread; res.bestTr = 0;
    ri.cSaEnum += (u32)ln.saEnum; ri.cNodes += (u32)ln.nodes; ri.cLeaves += (u32)ln.leaves;
    if (ln.overflow) {   // a cap of this tier was hit: flag the read for the next tier
        ri.flags |= 1 | (ln.overflow << 8);   // bits 8..: reason (analysis only)
        ri.cSaEnum = 0; ri.cNodes = 0; ri.cLeaves = 0;
        info[i] = ri;
        results[i] = res;
        return;
    }
    u32 nWfinal = ln.nW1;
    const int bestScore = ln.bestScore;
    if (mapMarker == 0 && bestScore == 0) { mapMarker = STAR_MARKER_NO_GOOD_WINDOW; }   // stitchPieces.cpp:344-348
    if (bestScore == 0) nWfinal = 0;
    u32 bestNMM = 0, bestNMatch = 0;
    if (ln.bestPool >= 0) {
        const TrHead& b = ln.pool[ln.bestPool].h;
        bestNMM = b.nMM; bestNMatch = b.nMatch; bestRLength = b.rLength;
    }
    u32 nTr = 0;
    #pragma unroll 1
    for (u32 w = 0; w < nWfinal; w++) {
        const u16* wt = ln.trPtr + ln.winBase[w];
        #pragma unroll 1
        for (u32 iTr = 0; iTr < ln.winN[w]; iTr++) {
            if (ln.pool[wt[iTr]].h.maxScore + P.outFilterMultimapScoreRange >= bestScore) nTr++;
        }
    }
    int unmapType = -1;
    if (nWfinal == 0) {
        unmapType = 0;
    } else if ((bestScore < P.outFilterScoreMin) || (bestScore < (int)(P.outFilterScoreMinOverLread * (double)(Lread - 1)))
               || (bestNMatch < P.outFilterMatchNmin) || (bestNMatch < (u64)(P.outFilterMatchNminOverLread * (double)(Lread - 1)))) {
        unmapType = 1;
    } else if ((bestNMM > ri.outFilterMismatchNmaxTotal) || (double(bestNMM) / double(bestRLength) > P.outFilterMismatchNoverLmax)) {
        unmapType = 2;
    } else if (nTr > P.outFilterMultimapNmax) {
        unmapType = 3;
    }
    res.unmapType = unmapType; res.nTr = nTr; res.mapMarker = mapMarker; res.bestScore = bestScore; res.bestNMM = bestNMM; res.bestRLength = bestRLength;
    if (unmapType < 0) {
        // trMult order = window order then in-window rank (:26-44); primary flag rules (:56-91)
        star_align_t* o = staged + (u64)i * ln.caps.nOut;
        u32 kOut = 0;
        u32 bestK = 0;
        #pragma unroll 1
        for (u32 w = 0; w < nWfinal; w++) {
            const u16* wt = ln.trPtr + ln.winBase[w];
            #pragma unroll 1
            for (u32 iTr = 0; iTr < ln.winN[w]; iTr++) {
                DevTr& t = ln.pool[wt[iTr]];
                if (t.h.maxScore + P.outFilterMultimapScoreRange >= bestScore) {
                    t.h.primaryFlag = 0;
                    exportAlign(t, ln.win[w].Chr, ln.win[w].Str, ln.win[w].Str, Lread, ix, o + kOut);
                    if ((int)wt[iTr] == ln.bestPool) bestK = kOut;
                    kOut++;
                }
            }
        }
        if (nTr == 1) {
            o[0].primaryFlag = 1;
        } else {
            u32 nbest = 0;
            if (P.outSAMmultNmax != (u64)-1) {   // bring the best alignments to the top (:60-67)
                for (u32 itr = 0; itr < nTr; itr++) {
                    if (o[itr].maxScore == bestScore) {
                        if (itr != nbest) { star_align_t tmp = o[itr]; o[itr] = o[nbest]; o[nbest] = tmp; }
                        if (bestK == itr) bestK = nbest; else if (bestK == nbest) bestK = itr;
                        ++nbest;
                    }
                }
            }
            if (P.outSAMprimaryFlagAllBestScore) {
                #pragma unroll 1
                for (u32 itr = 0; itr < nTr; itr++) if (o[itr].maxScore == bestScore) o[itr].primaryFlag = 1;
            } else if (P.outSAMmultNmax != (u64)-1) {
                o[0].primaryFlag = 1;
            } else {
                o[bestK].primaryFlag = 1;
            }
        }
        res.nTrOut = nTr;
        res.bestTr = bestK;
    }
    results[i] = res;
    info[i] = ri;
}

// ------------------------------------------------------------------------------------------------------------------
// Heavy-read hand-over.  A read whose windows hold many seeds (an upper bound of its DFS size is sum_w 2^nWA_w) is not
// stitched by its lane: the lane exports the windows + seeds to a pool in HBM and the read is stitched by a whole warp in
// stitch_heavy_kernel.  Record layout (8-byte aligned): u32 nWin, u32 nSeedsTotal, then nWin x {u32 Chr, u16 nWA, u8 Str, u8 pad},
// then the seeds of the windows back to back (24 B each).
struct HeavyWin { u32 Chr; u16 nWA; u8 Str; u8 pad; };

__device__ bool exportHeavy(Lane& ln, u8* __restrict__ heavyPool, u64 heavyPoolBytes, unsigned long long* __restrict__ heavyBump, u64& offOut) {
    u32 nWin = 0, nSeeds = 0;
    #pragma unroll 1
    for (u32 w = 0; w < ln.nW; w++) if (ln.win[w].nWA > 0) { nWin++; nSeeds += ln.win[w].nWA; }
    u64 bytes = 8 + (u64)nWin * sizeof(HeavyWin) + (u64)nSeeds * sizeof(Seed);
    bytes = (bytes + 15) & ~15ULL;
    u64 off = atomicAdd(heavyBump, (unsigned long long)bytes);
    if (off + bytes > heavyPoolBytes) return false;
    u8* p = heavyPool + off;
    ((u32*)p)[0] = nWin; ((u32*)p)[1] = nSeeds;
    HeavyWin* hw = (HeavyWin*)(p + 8);
    Seed* sd = (Seed*)(p + 8 + (u64)nWin * sizeof(HeavyWin));
    u32 k = 0, q = 0;
    #pragma unroll 1
    for (u32 w = 0; w < ln.nW; w++) {
        const Window& W = ln.win[w];
        if (W.nWA == 0) continue;
        HeavyWin h; h.Chr = W.Chr; h.nWA = W.nWA; h.Str = W.Str; h.pad = 0;
        hw[k++] = h;
        const Seed* src = ln.wa + (u64)w * ln.caps.spw;
        #pragma unroll 1
        for (u32 a = 0; a < W.nWA; a++) sd[q++] = src[a];
    }
    offOut = off;
    return true;
}

// Warp-lockstep state machine.  Every lane owns one read at a time (persistent lanes, ticket counter) and advances it one
// unit of work per loop iteration; lanes of a warp that are in the same phase execute that phase's step together, so the
// dominant units (one stitchAlignToTranscript call per DFS node, one leaf finalisation) run with many active lanes.
// A read never waits for its warp neighbours.
enum { PH_FETCH = 0, PH_WIN, PH_FLANK, PH_ASSIGN, PH_NEXTWIN, PH_NODE, PH_LEAF, PH_SELECT, PH_DONE };


#ifndef STITCH_MIN_BLOCKS
#define STITCH_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(128, STITCH_MIN_BLOCKS) stitch_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                     ReadInfo* __restrict__ info, const Piece* __restrict__ pieces, u32 nReads,
                                                     const u32* __restrict__ readList, u32* __restrict__ counter, u8* __restrict__ arenas,
                                                     Caps caps, star_read_result_t* __restrict__ results, star_align_t* __restrict__ staged,
                                                     const u32* __restrict__ order, u32 smemStride, HeavyArgs hv) {
    // readList != NULL : overflow tier, ticket k -> read readList[k], piece slab k
    // order    != NULL : first tier, ticket k -> read order[k] (heaviest reads first), piece slab = read id
    extern __shared__ u8 smem[];
    u8* R0 = smem + (size_t)threadIdx.x * 2 * smemStride;
    u8* R2 = R0 + smemStride;
    const u32 lane = threadIdx.x & 31;
    const u32 warpBase = threadIdx.x & ~31u;
    Lane ln;
    // DFS state in per-thread local memory (L1-resident, interleaved across lanes) instead of the HBM arena
    DevTr curL, leafL;
    Frame stackL[STAR_UNDO_DEPTH];
    u8 phL[STAR_DFS_MAX_DEPTH + 4];
    ln.cur = &curL; ln.leaf = &leafL; ln.stack = stackL; ln.ph = phL;
    ln.ix = &ix; ln.P = &P; ln.R0 = R0; ln.R2 = R2; ln.R = R0; ln.caps = caps;
    ln.memo = nullptr; ln.memoMask = 0; ln.memoBase = 0; ln.memoHit = 0; ln.memoMiss = 0; ln.lastSeed = -1; ln.coop = 0;
    {
        u8* a = arenas + (u64)(blockIdx.x * blockDim.x + threadIdx.x) * caps.arenaBytes;
        ln.win = (Window*)a; a += (u64)caps.maxW * sizeof(Window);
        ln.wa = (Seed*)a; a += (u64)caps.maxW * caps.spw * sizeof(Seed);
        ln.pool = (DevTr*)a; a += (u64)caps.maxTr * sizeof(DevTr);
        a += 2 * sizeof(DevTr) + (u64)(caps.spw + 2) * sizeof(Frame);   // (layout kept; DFS state lives in local memory)
        ln.trPtr = (u16*)a; a += (u64)caps.maxTr * sizeof(u16);
        ln.winBase = (u16*)a; a += (u64)caps.maxW * sizeof(u16);
        ln.winN = (u16*)a;
    }
    u32 phase = PH_FETCH;
    u32 k = 0, i = 0;
    ReadInfo ri;
    const Piece* PC = nullptr;
    u32 nP = 0, iP = 0;
    u64 iSA = 0, iSAend = 0;
    Piece p;
    p.SAstart = 0; p.rStart = 0; p.Length = 0; p.Nrep = 0; p.Dir = 0; p.iFrag = 0;
    bool tooManyAnchors = false;
    u32 mapMarker = 0, iW = 0, Chr = 0, Str = 0, bestRLength = 0;
    u16* wTr = nullptr;
    u16 nWinTr = 0;

    long long pc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tPrev = clock64();
#define PHASE_TICK(slot) do { long long tn_ = clock64(); pc[slot] += tn_ - tPrev; tPrev = tn_; } while (0)
    for (;;) {
        // ------------------------------------------------------------------ fetch the next read
        bool needCopy = false;
        if (phase == PH_FETCH) {
            k = atomicAdd(counter, 1u);
            if (k >= nReads) {
                phase = PH_DONE;
            } else {
                i = readList ? readList[k] : (order ? order[k] : k);
                ri = info[i];
                readBegin(ln, ri);
                mapMarker = 0; bestRLength = 0;
                tooManyAnchors = false;
                if (ri.flags) {   // seed kernel overflowed (bit0) or hit the fatal piece limit (bit1): the read is redone / reported elsewhere
                    star_read_result_t res;
                    res.unmapType = 0; res.nTr = 0; res.nTrOut = 0; res.mapMarker = 0; res.trOffset = 0; res.bestScore = 0; res.bestNMM = 0;
                    res.bestRLength = 0; res.Lread = ri.Lread; res.bestTr = 0;
                    results[i] = res;
                } else if (ri.Lread < P.outFilterMatchNmin) {   // ReadAlign_mapOneRead.cpp:100-115
                    mapMarker = STAR_MARKER_READ_TOO_SHORT; bestRLength = 0; phase = PH_SELECT;
                } else if (ri.Nsplit == 0) {
                    mapMarker = STAR_MARKER_NO_GOOD_PIECES; bestRLength = ri.split1_0; phase = PH_SELECT;
                } else if (ri.nA == 0) {
                    mapMarker = STAR_MARKER_ALL_PIECES_EXCEED_seedMultimapNmax; bestRLength = ri.multNminL; phase = PH_SELECT;
                } else {
                    needCopy = true;
                    PC = pieces + (u64)(readList ? k : i) * caps.maxP;
                    nP = ri.nP; iP = 0; iSA = 0; iSAend = 0;
                    ln.nW = 0;
                    phase = PH_WIN;
                }
            }
        }
        {   // the whole warp copies the reads of the lanes that just fetched one (coalesced) into their shared-memory rows
            u32 m = __ballot_sync(0xffffffffu, needCopy);
            #pragma unroll 1
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                u32 ri_i = __shfl_sync(0xffffffffu, i, src);
                u32 L = __shfl_sync(0xffffffffu, ln.Lread, src);
                const u8* g = reads + (u64)ri_i * stride;
                u8* d0 = smem + (size_t)(warpBase + src) * 2 * smemStride;
                u8* d2 = d0 + smemStride;
                #pragma unroll 1
                for (u32 b = lane; b < L; b += 32) {
                    u8 c = g[b];
                    d0[b] = c;
                    d2[L - 1 - b] = c < 4 ? 3 - c : c;
                }
            }
            __syncwarp();
        }
        PHASE_TICK(0);
        // ------------------------------------------------------------------ window creation, one SA locus per step (:41-93)
        if (phase == PH_WIN) {
            if (iSA >= iSAend) {
                #pragma unroll 1
                while (iP < nP) {
                    p = PC[iP];
                    iP++;
                    if (p.Nrep <= P.winAnchorMultimapNmax) { iSA = p.SAstart; iSAend = p.SAstart + p.Nrep; break; }
                }
                if (iSA >= iSAend) phase = PH_FLANK;
            }
            if (phase == PH_WIN) {
                u64 raw[LOCI_PER_STEP];
                u32 nl = (u32)(iSAend - iSA < LOCI_PER_STEP ? iSAend - iSA : LOCI_PER_STEP);
#pragma unroll
                for (u32 q = 0; q < LOCI_PER_STEP; q++) if (q < nl) raw[q] = packedGet(ix.SA, ix.saBits, iSA + q);
                const u64 aLength = p.Length;
#pragma unroll
                for (u32 q = 0; q < LOCI_PER_STEP; q++) {
                    if (q >= nl || ln.overflow) break;
                    ln.saEnum++;
                    iSA++;
                    u64 a1 = raw[q];
                    u32 aStr = (u32)(a1 >> ix.GstrandBit);
                    a1 &= ix.GstrandMask;
                    if (p.Dir == 1 && aStr == 0) { aStr = 1; }
                    else if (p.Dir == 0 && aStr == 1) { a1 = ix.nGenome - (aLength + a1); }
                    else if (p.Dir == 1 && aStr == 1) { aStr = 0; a1 = ix.nGenome - (aLength + a1); }
                    bool stopPiece = false;
                    if (a1 >= ix.sjGstart) {
                        u64 a1D, aLengthD, a1A, aLengthA; u32 sj1;
                        if (sjAlignSplit(ix, a1, aLength, a1D, aLengthD, a1A, aLengthA, sj1)) {
                            if (createExtendWindowsWithAlign(ln, a1D, aStr) == 101) stopPiece = true;
                            else if (createExtendWindowsWithAlign(ln, a1A, aStr) == 101) stopPiece = true;
                        }
                    } else {
                        if (createExtendWindowsWithAlign(ln, a1, aStr) == 101) stopPiece = true;
                    }
                    if (stopPiece) { iSA = iSAend; break; }   // EXIT_createExtendWindowsWithAlign_TOO_MANY_WINDOWS breaks the loop of this piece only
                }
                if (ln.overflow) phase = PH_SELECT;
            }
        }
        PHASE_TICK(1);
        // ------------------------------------------------------------------ flanks (:96-118)
        if (phase == PH_FLANK) {
            #pragma unroll 1
            for (u32 iWin = 0; iWin < ln.nW; iWin++) {
                Window& W = ln.win[iWin];
                if (W.gStart <= W.gEnd) {
                    u64 wb = W.gStart;
                    #pragma unroll 1
                    for (u64 ii = 0; ii < P.winFlankNbins && wb > 0 && chrOfBin(ln, wb - 1) == W.Chr; ii++) wb--;
                    W.gStart = (u32)wb;
                    wb = W.gEnd;
                    #pragma unroll 1
                    for (u64 ii = 0; ii < P.winFlankNbins && wb + 1 < P.winBinN && chrOfBin(ln, wb + 1) == W.Chr; ii++) wb++;
                    W.gEnd = (u32)wb;
                }
                W.nWA = 0; W.WALrec = 0;
            }
            iP = 0; iSA = 0; iSAend = 0;
            phase = PH_ASSIGN;
        }
        PHASE_TICK(2);
        // ------------------------------------------------------------------ seed -> window assignment, one SA locus per step (:129-185)
        if (phase == PH_ASSIGN) {
            if (iSA >= iSAend) {
                if (iP < nP) { p = PC[iP]; iP++; iSA = p.SAstart; iSAend = p.SAstart + p.Nrep; }
                else {
                    // init per-window stitching (:262-270)
                    if (tooManyAnchors) ln.nW = 0;   // assignAlignToWindow.cpp:77-81; ends as MARKER_NO_GOOD_WINDOW
                    bool heavy = false;
                    if (hv.estLimit) {
                        u64 est = 0;
                        #pragma unroll 1
                        for (u32 w = 0; w < ln.nW; w++) { u32 a = ln.win[w].nWA; if (a) est += 1ULL << (a < 20 ? a : 20); }
                        heavy = est > hv.estLimit;
                    }
                    if (heavy) {
                        u64 off;
                        if (exportHeavy(ln, hv.pool, hv.poolBytes, hv.bump, off)) {
                            hv.readOff[i] = off;
                            u32 q = atomicAdd(hv.count, 1u);
                            hv.list[q] = i;
                            ri.cSaEnum = (u32)ln.saEnum; ri.cNodes = 0; ri.cLeaves = 0;   // the heavy kernel adds its DFS work
                            info[i] = ri;
                            phase = PH_FETCH;
                        } else {
                            ln.overflow = 4;      // reason 4: export pool exhausted: next tier
                            phase = PH_SELECT;
                        }
                    } else {
                        #pragma unroll 1
                        for (u32 q = 0; q < caps.maxTr; q++) ln.trPtr[q] = (u16)q;
                        iW = 0;
                        phase = PH_NEXTWIN;
                    }
                }
            }
            if (phase == PH_ASSIGN) {
                u64 raw[LOCI_PER_STEP];
                u32 nl = (u32)(iSAend - iSA < LOCI_PER_STEP ? iSAend - iSA : LOCI_PER_STEP);
#pragma unroll
                for (u32 q = 0; q < LOCI_PER_STEP; q++) if (q < nl) raw[q] = packedGet(ix.SA, ix.saBits, iSA + q);
                const u64 aNrep = p.Nrep, aLength = p.Length;
                const u32 aFrag = p.iFrag;
                const bool aAnchor = aNrep <= P.winAnchorMultimapNmax;
                const u32 Lread = ln.Lread;
#pragma unroll
                for (u32 q = 0; q < LOCI_PER_STEP; q++) {
                    if (q >= nl) break;
                    ln.saEnum++;
                    iSA++;
                    u64 a1 = raw[q];
                    u32 aStr = (u32)(a1 >> ix.GstrandBit);
                    a1 &= ix.GstrandMask;
                    u64 aRstart = p.rStart;
                    if (p.Dir == 1 && aStr == 0) { aStr = 1; aRstart = Lread - (aLength + aRstart); }
                    else if (p.Dir == 0 && aStr == 1) { aRstart = Lread - (aLength + aRstart); a1 = ix.nGenome - (aLength + a1); }
                    else if (p.Dir == 1 && aStr == 1) { aStr = 0; a1 = ix.nGenome - (aLength + a1); }
                    if (a1 >= ix.sjGstart) {
                        u64 a1D, aLengthD, a1A, aLengthA; u32 isj1;
                        if (sjAlignSplit(ix, a1, aLength, a1D, aLengthD, a1A, aLengthA, isj1)) {
                            if (!assignAlignToWindow(ln, a1D, aLengthD, aStr, aNrep, aFrag, aRstart, aAnchor, isj1)) tooManyAnchors = true;
                            else if (!assignAlignToWindow(ln, a1A, aLengthA, aStr, aNrep, aFrag, aRstart + aLengthD, aAnchor, isj1)) tooManyAnchors = true;
                        }
                    } else {
                        if (!assignAlignToWindow(ln, a1, aLength, aStr, aNrep, aFrag, aRstart, aAnchor, SJA_NONE)) tooManyAnchors = true;
                    }
                    if (tooManyAnchors) { iSA = iSAend; iP = nP; break; }   // the rest of the reference's enumeration is dead work (nW=0)
                }
            }
        }
        PHASE_TICK(3);
        // ------------------------------------------------------------------ next window with seeds (:268-299)
        if (phase == PH_NEXTWIN) {
            #pragma unroll 1
            while (iW < ln.nW && ln.win[iW].nWA == 0) iW++;
            int rc = iW < ln.nW ? windowBegin(ln, wTr, nWinTr) : 1;
            if (rc == 2) { ln.overflow = 3; phase = PH_SELECT; }   // reason 3: transcript pool
            else if (rc == 1) { phase = PH_SELECT; }
            else {
                Chr = ln.win[iW].Chr; Str = ln.win[iW].Str;
                ln.R = Str == 0 ? R0 : R2;
                dfsInit(ln);
                phase = PH_NODE;
            }
        }
        PHASE_TICK(4);
        // ------------------------------------------------------------------ one DFS unit: one seed-include attempt
        if (phase == PH_NODE) {
            int r = dfsStep(ln, ln.wa + (u64)iW * caps.spw, ln.win[iW].nWA);
            if (r == DFS_LEAF) {
                phase = PH_LEAF;
            } else if (r == DFS_DONE) {
                windowEnd(ln, Chr, Str, wTr, nWinTr);
                iW++;
                phase = PH_NEXTWIN;
            }
        }
        PHASE_TICK(5);
        // ------------------------------------------------------------------ leaf: extend, filter, score, record (:19-306)
        if (phase == PH_LEAF) {
            ln.leaves++;
            if (nWinTr > caps.maxTr - ln.trNtotal - 1) {   // pool of this lane is full: next tier
                ln.overflow = 3;
                phase = PH_SELECT;
            } else {
                if (evalLeaf(ln, ln.leafScore, ln.leafR2, ln.leafG2, Chr, Str, Str)) recordLeaf(ln, wTr, &nWinTr);
                phase = PH_NODE;
            }
        }
        PHASE_TICK(6);
        // ------------------------------------------------------------------ multMapSelect, mappedFilter, export
        if (phase == PH_SELECT) {
            selectExport(ln, ri, i, mapMarker, bestRLength, results, staged, info);
            phase = PH_FETCH;
        }
        PHASE_TICK(7);
        if (__all_sync(0xffffffffu, phase == PH_DONE)) break;
    }
    #pragma unroll 1
    for (int q = 0; q < 8; q++) PROF_ADD(q, pc[q]);
}

// ------------------------------------------------------------------------------------------------------------------
// Heavy reads: ONE WARP per read.  The DFS of every window is cut into prefix sub-trees (the first d include/exclude decisions
// fixed; task id ascending = the reference's DFS order, include before exclude).  Lanes pull tasks from a warp-wide ticket and
// evaluate them independently: the expensive part of a leaf (stitch chain, end extension, filters, score: evalLeaf) is a pure
// function of the path, so it parallelises exactly.  Each surviving leaf is stored as a 16-byte candidate {include mask, score,
// iFrag}.  Then lane 0 replays the reference's order-dependent part (recordLeaf: maxScoreMate, record test, blocksOverlap dedup,
// ordered insert) over the candidates in window order / task order / leaf order, re-materialising the transcript (replayPath +
// evalLeaf) only for the few candidates that pass the record test.  Result: identical to the sequential recursion.
// trOff: offset (in 8-byte words) of the evaluated transcript stored by the lane for candidates that are likely to pass the record
// test (saves the replay in the R phase); 0xFFFFFFFF = not stored, the R phase replays the path.
#define CAND_PER_BLOCK 31
struct CandBlock { u32 next; u32 count; Cand c[CAND_PER_BLOCK]; };   // 8 + 31*16 = 504 bytes
struct TaskOut { u32 first, last; };                                 // candidate blocks of a task (0xFFFFFFFF = none)


__device__ bool replayPath(Lane& ln, const Seed* __restrict__ WA, u32 nA, u64 mask, int& Score, u32& tR2, u64& tG2) {
    DevTr* t = ln.cur;
    dfsInit(ln);
    Score = 0; tR2 = 0; tG2 = 0;
    #pragma unroll 1
    for (u32 iA = 0; iA < nA; iA++) {
        if (!((mask >> iA) & 1ULL)) continue;
        const Seed s = WA[iA];
        int dScore;
        if (t->h.nExons > 0) {
            bool hit;
            dScore = stitchMemoized(ln, tR2, tG2, s, iA, t, hit);
        } else {
            t->ex[0].R = s.rStart; t->h.rStart = s.rStart;
            t->ex[0].G = s.gStart; t->h.gStart = s.gStart;
            t->ex[0].L = s.Length; t->ex[0].iFrag = s.iFrag; t->ex[0].sjA = s.sjA;
            t->ex[0].canon = 0; t->ex[0].annot = 0; t->ex[0].sjStr = 0; t->ex[0].shL = 0; t->ex[0].shR = 0;
            t->h.nExons = 1;
            dScore = s.Length;
            t->h.nMatch = s.Length;
        }
        if (dScore <= -1000000) return false;
        ln.lastSeed = (int)iA;
        if (s.Nrep == 1) t->h.nUnique++;
        if (s.Anchor > 0) t->h.nAnchor++;
        Score += dScore;
        tR2 = (u32)s.rStart + s.Length - 1;
        tG2 = s.gStart + s.Length - 1;
    }
    return tR2 != 0;
}

// ---- warp-cooperative window creation / seed assignment (same semantics as createExtendWindowsWithAlign / assignAlignToWindow
// above; all 32 lanes call these with IDENTICAL arguments, the window table lives in shared memory, the seeds of window w at
// wa[w*spw ..] in the warp's arena).  Lanes scan windows / seeds 32 at a time and agree through ballots and shuffles.
struct WarpWin {
    Window* swin;     // shared memory, caps.maxW entries
    Seed* wa;         // arena of the warp
    u32 nW;
    u32 spw;
    u32 lane;
};

__device__ __forceinline__ u64 warpMaxU64(u64 v) {
    #pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) { u64 x = __shfl_xor_sync(0xffffffffu, v, o); v = x > v ? x : v; }
    return v;
}
__device__ __forceinline__ u64 warpMinU64(u64 v) {
    #pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) { u64 x = __shfl_xor_sync(0xffffffffu, v, o); v = x < v ? x : v; }
    return v;
}

// ReadAlign_createExtendWindowsWithAlign.cpp:7-84, cooperative.  Returns 0, 101 (TOO_MANY_WINDOWS) or 102 (tier cap: overflow).
__device__ int coopCreateWindow(const Lane& ln, WarpWin& ww, u64 a1, u32 aStr) {
    const star_params_t& P = *ln.P;
    const u64 aBin = a1 >> P.winBinNbits;
    const u64 lo = aBin > P.winAnchorDistNbins ? aBin - P.winAnchorDistNbins : 0;
    const u64 hiX = aBin + P.winAnchorDistNbins + 1 < P.winBinN ? aBin + P.winAnchorDistNbins + 1 : P.winBinN;
    bool owned = false;
    u64 candL = 0;                      // (gEnd+1)<<16 | w   (0 = none); max wins
    u64 candR = ~0ULL;                  // gStart<<16 | w     (~0 = none); min wins
    #pragma unroll 1
    for (u32 w = ww.lane; w < ww.nW; w += 32) {
        const Window W = ww.swin[w];
        if (W.Str != aStr || W.gStart > W.gEnd) continue;
        const u64 s0 = W.gStart, e0 = W.gEnd;
        if (s0 <= aBin && aBin <= e0) owned = true;
        if (aBin > 0 && e0 < aBin && e0 >= lo) { u64 c = ((e0 + 1) << 16) | w; if (c > candL) candL = c; }
        if (aBin + 1 < P.winBinN && s0 > aBin && s0 < hiX) { u64 c = (s0 << 16) | w; if (c < candR) candR = c; }
    }
    if (__any_sync(0xffffffffu, owned)) return 0;
    candL = warpMaxU64(candL);
    candR = warpMinU64(candR);
    const int left = candL ? (int)(candL & 0xffff) : -1;
    const int right = candR != ~0ULL ? (int)(candR & 0xffff) : -1;
    const u64 bestL = candL ? (candL >> 16) - 1 : 0, bestR = candR != ~0ULL ? candR >> 16 : 0;
    const bool flagMergeLeft = left >= 0 && chrOfBin(ln, bestL) == chrOfBin(ln, aBin);
    const bool flagMergeRight = right >= 0 && chrOfBin(ln, bestR) == chrOfBin(ln, aBin);
    int rc = 0;
    if (!flagMergeLeft && !flagMergeRight) {
        if (ww.nW >= ln.caps.maxW) return 102;
        if (ww.lane == 0) {
            Window nw;
            nw.gStart = (u32)aBin; nw.gEnd = (u32)aBin; nw.Chr = chrOfBin(ln, aBin); nw.nWA = 0; nw.WALrec = 0; nw.Str = (u8)aStr;
            nw.pad[0] = nw.pad[1] = nw.pad[2] = 0;
            ww.swin[ww.nW] = nw;
        }
        ww.nW++;
        if (ww.nW >= P.alignWindowsPerReadNmax) { ww.nW = (u32)P.alignWindowsPerReadNmax - 1; rc = 101; }
    } else {
        if (ww.lane == 0) {
            u64 iBinLeft = aBin, iBinRight = aBin;
            int iWin = -1;
            if (flagMergeLeft) { iWin = left; iBinLeft = ww.swin[left].gStart; }
            if (flagMergeRight) { iBinRight = ww.swin[right].gEnd; if (!flagMergeLeft) iWin = right; }
            ww.swin[iWin].gStart = (u32)iBinLeft;
            ww.swin[iWin].gEnd = (u32)iBinRight;
            if (flagMergeLeft && flagMergeRight) { ww.swin[right].gStart = 1; ww.swin[right].gEnd = 0; }
        }
    }
    __syncwarp();
    return rc;
}

// ReadAlign_assignAlignToWindow.cpp:6-130, cooperative, window iW already looked up.  Returns false on MARKER_TOO_MANY_ANCHORS.
__device__ bool coopAssign(const Lane& ln, WarpWin& ww, int iW, u64 a1, u64 aLength, u64 aNrep, u32 aFrag, u64 aRstart, bool aAnchor, u32 sjA) {
    const star_params_t& P = *ln.P;
    if (iW < 0) return true;
    Window* Wp = &ww.swin[iW];
    const u32 WALrec0 = Wp->WALrec;
    if (!aAnchor && aLength < WALrec0) return true;
    Seed* WA = ww.wa + (u64)iW * ww.spw;
    u32 nWA = Wp->nWA;
    const u32 lane = ww.lane;
    // my (up to two) seeds: j0 = lane, j1 = lane + 32
    Seed e0, e1;
    const bool v0 = lane < nWA, v1 = lane + 32 < nWA;
    if (v0) e0 = WA[lane];
    if (v1) e1 = WA[lane + 32];
    auto overlaps = [&](const Seed& s0) {
        return aFrag == s0.iFrag && s0.sjA == sjA && a1 + s0.rStart == s0.gStart + aRstart
               && ((aRstart >= s0.rStart && aRstart < (u64)s0.rStart + s0.Length) || (aRstart + aLength >= s0.rStart && aRstart + aLength < (u64)s0.rStart + s0.Length));
    };
    Seed nsd;
    setSeed(nsd, a1, aLength, aNrep, aFrag, aRstart, aAnchor, sjA);
    {
        u32 m0 = __ballot_sync(0xffffffffu, v0 && overlaps(e0));
        u32 m1 = __ballot_sync(0xffffffffu, v1 && overlaps(e1));
        int iA = m0 ? __ffs(m0) - 1 : (m1 ? 32 + __ffs(m1) - 1 : -1);
        if (iA >= 0) {
            u32 LiA = __shfl_sync(0xffffffffu, iA < 32 ? (u32)e0.Length : (u32)e1.Length, iA & 31);
            if (aLength > LiA) {
                // insertion point: first iA0 != iA with aRstart < rStart, else nWA
                u32 b0 = __ballot_sync(0xffffffffu, v0 && (int)lane != iA && aRstart < e0.rStart);
                u32 b1 = __ballot_sync(0xffffffffu, v1 && (int)(lane + 32) != iA && aRstart < e1.rStart);
                int iA0 = b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : (int)nWA);
                if (iA0 > iA) --iA0;
                __syncwarp();
                if (iA0 < iA) {   // elements iA0..iA-1 move up by one
                    if (v0 && (int)lane >= iA0 && (int)lane < iA) WA[lane + 1] = e0;
                    if (v1 && (int)(lane + 32) >= iA0 && (int)(lane + 32) < iA) WA[lane + 33] = e1;
                } else if (iA0 > iA) {   // elements iA+1..iA0 move down by one
                    if (v0 && (int)lane > iA && (int)lane <= iA0) WA[lane - 1] = e0;
                    if (v1 && (int)(lane + 32) > iA && (int)(lane + 32) <= iA0) WA[lane + 31] = e1;
                }
                __syncwarp();
                if (lane == 0) WA[iA0] = nsd;
                __syncwarp();
            }
            return true;
        }
    }
    if (nWA == P.seedPerWindowNmax) {
        u32 rec = ln.Lread + 1;
        if (v0 && e0.Anchor != 1 && e0.Length < rec) rec = e0.Length;
        if (v1 && e1.Anchor != 1 && e1.Length < rec) rec = e1.Length;
        rec = (u32)warpMinU64(rec);
        if (lane == 0) Wp->WALrec = (u16)rec;
        __syncwarp();
        if (rec == ln.Lread + 1) return false;
        if (!aAnchor && aLength < rec) return true;
        const bool k0 = v0 && (e0.Anchor == 1 || e0.Length > rec), k1 = v1 && (e1.Anchor == 1 || e1.Length > rec);
        const u32 km0 = __ballot_sync(0xffffffffu, k0), km1 = __ballot_sync(0xffffffffu, k1);
        const u32 below = (1u << lane) - 1;
        __syncwarp();
        if (k0) WA[__popc(km0 & below)] = e0;
        if (k1) WA[__popc(km0) + __popc(km1 & below)] = e1;
        __syncwarp();
        nWA = __popc(km0) + __popc(km1);
        if (lane == 0) Wp->nWA = (u16)nWA;
        __syncwarp();
    }
    const u32 WALrec = Wp->WALrec;
    if (aAnchor || aLength > WALrec) {
        // reload (the purge may have moved the seeds)
        const bool w0 = lane < nWA, w1 = lane + 32 < nWA;
        if (w0) e0 = WA[lane];
        if (w1) e1 = WA[lane + 32];
        u32 b0 = __ballot_sync(0xffffffffu, w0 && aRstart < e0.rStart);
        u32 b1 = __ballot_sync(0xffffffffu, w1 && aRstart < e1.rStart);
        int iA = b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : (int)nWA);
        __syncwarp();
        if (w0 && (int)lane >= iA) WA[lane + 1] = e0;
        if (w1 && (int)(lane + 32) >= iA) WA[lane + 33] = e1;
        __syncwarp();
        if (lane == 0) { WA[iA] = nsd; Wp->nWA = (u16)(nWA + 1); }
        __syncwarp();
    }
    return true;
}

// Window phases of a heavy read, executed by the whole warp (all lanes call with identical arguments): mode A imports the windows +
// seeds a lane of stitch_kernel exported, mode B builds them cooperatively from the stored pieces.  Result: window table in shared
// memory (swin[0..nWin)), seeds of window w at ln.wa[w*caps.spw ..].
__device__ __forceinline__ void warpBuildWindows(Lane& ln, WarpWin& ww, const DevIndex& ix, const star_params_t& P, const ReadInfo& ri, u32 i, u32 slab,
                                                 const Piece* __restrict__ pieces, const u64* __restrict__ heavyOff, const u8* __restrict__ heavyPool,
                                                 const Caps& caps, Window* swin, u32 lane, u32& nWin, u32& overReason, u32* sortScratch = nullptr,
                                                 u32* binFilter = nullptr, u32 filterLog2 = 0) {
    const u32 Lread = ri.Lread;
    bool tooManyAnchors = false;
    u64 saEnum = 0;
    if (heavyPool) {
        // ---- mode A: import the exported windows + seeds
        const u8* rec = heavyPool + heavyOff[i];
        nWin = ((const u32*)rec)[0];
        const HeavyWin* hw = (const HeavyWin*)(rec + 8);
        const Seed* seeds = (const Seed*)(rec + 8 + (u64)nWin * sizeof(HeavyWin));
        if (nWin > caps.maxW) { overReason = 5; nWin = 0; }
        u32 sd = 0;
        #pragma unroll 1
        for (u32 w = 0; w < nWin; w++) {   // uniform loop; lanes copy the seeds of window w
            const HeavyWin h = hw[w];
            if (lane == 0) {
                Window nw; nw.gStart = 0; nw.gEnd = 0; nw.Chr = h.Chr; nw.nWA = h.nWA; nw.WALrec = 0; nw.Str = h.Str; nw.pad[0] = nw.pad[1] = nw.pad[2] = 0;
                swin[w] = nw;
            }
            #pragma unroll 1
            for (u32 a = lane; a < h.nWA; a += 32) ln.wa[(u64)w * caps.spw + a] = seeds[sd + a];
            sd += h.nWA;
        }
        ww.nW = nWin;
    } else {
        // ---- mode B: cooperative window creation and seed assignment (ReadAlign_stitchPieces.cpp:41-185)
        const Piece* PC = pieces + (u64)slab * caps.maxP;   // caps.maxP = slab stride of the seed kernel of this tier; slab = read id, or list position in a tier
        const u32 nP = ri.nP;
        ww.nW = 0;
        #pragma unroll 1
        for (u32 iP = 0; iP < nP && !overReason; iP++) {
            const Piece p = PC[iP];
            if (p.Nrep > P.winAnchorMultimapNmax) continue;
            const u64 aLength = p.Length;
            bool stopPiece = false;
            #pragma unroll 1
            for (u64 base = 0; base < p.Nrep && !stopPiece && !overReason; base += 32) {
                const u32 nl = (u32)(p.Nrep - base < 32 ? p.Nrep - base : 32);
                u64 a1 = 0, a1A = 0; u32 aStr = 0; u32 kind = 0;   // kind: 0 skip, 1 genomic, 2 sjdb (donor a1, acceptor a1A)
                if (lane < nl) {
                    a1 = packedGet(ix.SA, ix.saBits, p.SAstart + base + lane);
                    aStr = (u32)(a1 >> ix.GstrandBit);
                    a1 &= ix.GstrandMask;
                    if (p.Dir == 1 && aStr == 0) { aStr = 1; }
                    else if (p.Dir == 0 && aStr == 1) { a1 = ix.nGenome - (aLength + a1); }
                    else if (p.Dir == 1 && aStr == 1) { aStr = 0; a1 = ix.nGenome - (aLength + a1); }
                    kind = 1;
                    if (a1 >= ix.sjGstart) {
                        u64 a1D, aLengthD, aLengthA; u32 sj1;
                        if (sjAlignSplit(ix, a1, aLength, a1D, aLengthD, a1A, aLengthA, sj1)) { a1 = a1D; kind = 2; } else kind = 0;
                    }
                }
                #pragma unroll 1
                for (u32 q = 0; q < nl; q++) {
                    saEnum++;
                    const u32 kq = __shfl_sync(0xffffffffu, kind, q);
                    const u64 a1q = __shfl_sync(0xffffffffu, a1, q);
                    const u64 a1Aq = __shfl_sync(0xffffffffu, a1A, q);
                    const u32 sq = __shfl_sync(0xffffffffu, aStr, q);
                    if (kq == 0) continue;
                    int rc = coopCreateWindow(ln, ww, a1q, sq);
                    if (rc == 0 && kq == 2) rc = coopCreateWindow(ln, ww, a1Aq, sq);
                    if (rc == 102) { overReason = 1; break; }
                    if (rc == 101) { stopPiece = true; break; }
                }
            }
        }
        // flanks :96-118 (one window per lane)
        for (u32 w = lane; w < ww.nW; w += 32) {
            Window W = swin[w];
            if (W.gStart <= W.gEnd) {
                u64 wb = W.gStart;
                #pragma unroll 1
                for (u64 ii = 0; ii < P.winFlankNbins && wb > 0 && chrOfBin(ln, wb - 1) == W.Chr; ii++) wb--;
                W.gStart = (u32)wb;
                wb = W.gEnd;
                #pragma unroll 1
                for (u64 ii = 0; ii < P.winFlankNbins && wb + 1 < P.winBinN && chrOfBin(ln, wb + 1) == W.Chr; ii++) wb++;
                W.gEnd = (u32)wb;
            }
            W.nWA = 0; W.WALrec = 0;
            swin[w] = W;
        }
        __syncwarp();
        // Reads with many windows (hundreds for long reads with many short multi-mapping seeds): the window of a locus is found by bisection
        // over the live windows sorted by (strand, first bin) instead of a scan of the whole table per locus.  Live windows of one strand
        // do not overlap (they were created more than winAnchorDistNbins apart and the flanks add 2 x winFlankNbins <= that), so the
        // bisection finds the window the scan finds; should two neighbours in that order both hold the bin, the lane falls back to the scan.
        const bool sortedLookup = sortScratch != nullptr && ww.nW > caps.sortMinW && ww.nW <= 4096 && P.winBinN < (1ULL << 19);
        u32 nLive = 0;
        if (sortedLookup) {
            u32* tmpKeys = (u32*)ln.win;   // (the arena's window array is not used by the cooperative path: scratch for the unsorted keys)
            #pragma unroll 1
            for (u32 w = lane; w < ww.nW; w += 32) {
                const Window W = swin[w];
                tmpKeys[w] = W.gStart <= W.gEnd ? ((((u32)W.Str << 19) | W.gStart) << 12) | w : 0xFFFFFFFFu;
            }
            __syncwarp();
            #pragma unroll 1
            for (u32 w = lane; w < ww.nW; w += 32) {
                const u32 mine = tmpKeys[w];
                u32 rank = 0;
                #pragma unroll 1
                for (u32 v = 0; v < ww.nW; v++) { const u32 o = tmpKeys[v]; rank += (o < mine) || (o == mine && v < w); }
                sortScratch[rank] = mine;
            }
            __syncwarp();
            u32 live = 0;
            #pragma unroll 1
            for (u32 w = lane; w < ww.nW; w += 32) live += tmpKeys[w] != 0xFFFFFFFFu;
            for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
            nLive = live;
        }
        // Almost every locus of a multi-mapping piece falls into no window at all.  One bit per (strand, bin) covered by a live window, hashed
        // into 2^filterLog2 bits of shared memory, answers "in no window" with one load; only loci that pass are looked up (a bit set by
        // another bin costs a lookup that finds nothing: the result does not depend on the filter).
        bool useFilter = binFilter != nullptr && filterLog2 >= 8 && caps.binFilter != 0 && ww.nW > 0;
        auto filterHash = [&](u32 aStr, u64 bin) -> u32 { return (((u32)bin * 2654435761u) ^ (aStr * 0x7F4A7C15u)) >> (32 - filterLog2); };
        if (useFilter) {
            u32 covered = 0;
            #pragma unroll 1
            for (u32 w = lane; w < ww.nW; w += 32) { const Window W = swin[w]; if (W.gStart <= W.gEnd) covered += W.gEnd - W.gStart + 1; }
            for (int o = 16; o > 0; o >>= 1) covered += __shfl_xor_sync(0xffffffffu, covered, o);
            if ((u64)covered * 4 > (1ULL << filterLog2)) useFilter = false;   // too many bins for this many bits (most loci would pass)
        }
        if (useFilter) {
            #pragma unroll 1
            for (u32 q = lane; q < (1u << filterLog2) / 32; q += 32) binFilter[q] = 0;
            __syncwarp();
            #pragma unroll 1
            for (u32 w = lane; w < ww.nW; w += 32) {
                const Window W = swin[w];
                if (W.gStart > W.gEnd) continue;
                #pragma unroll 1
                for (u32 b = W.gStart; b <= W.gEnd; b++) { const u32 h = filterHash(W.Str, b); atomicOr(&binFilter[h >> 5], 1u << (h & 31)); }
            }
            __syncwarp();
        }
        auto findWindow = [&](u32 aStr, u64 bin) -> int {   // index of the live window of strand aStr that holds `bin`, -1 if none
            if (nLive == 0 || bin >= (1ULL << 19)) return -1;
            const u32 target = ((((u32)aStr << 19) | (u32)bin) << 12) | 0xFFFu;
            u32 lo = 0, hi = nLive;                           // first sorted entry > target
            #pragma unroll 1
            while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (sortScratch[mid] <= target) lo = mid + 1; else hi = mid; }
            if (lo == 0) return -1;
            const u32 w = sortScratch[lo - 1] & 0xFFFu;
            const Window W = swin[w];
            if (W.Str != aStr || bin > W.gEnd) return -1;     // (gStart <= bin by the order)
            if (lo >= 2) {                                    // the window in front must not hold the bin as well
                const Window V = swin[sortScratch[lo - 2] & 0xFFFu];
                if (V.Str == aStr && V.gEnd >= bin) return -2;
            }
            return (int)w;
        };
        // assignment :129-185
        #pragma unroll 1
        for (u32 iP = 0; iP < nP && !overReason && !tooManyAnchors; iP++) {
            const Piece p = PC[iP];
            const u64 aNrep = p.Nrep, aLength = p.Length;
            const u32 aFrag = p.iFrag;
            const bool aAnchor = aNrep <= P.winAnchorMultimapNmax;
            u64 saNext = lane < p.Nrep ? packedGet(ix.SA, ix.saBits, p.SAstart + lane) : 0;
            #pragma unroll 1
            for (u64 base = 0; base < p.Nrep && !tooManyAnchors; base += 32) {
                const u32 nl = (u32)(p.Nrep - base < 32 ? p.Nrep - base : 32);
                u64 a1 = 0, a1A = 0, aRstart = 0, aLengthD = 0, aLengthA = 0; u32 aStr = 0, kind = 0, isj = SJA_NONE;
                int wD = -1, wA = -1;
                const u64 saCur = saNext;   // the SA rows of the next 32 loci are in flight while these are looked up (multi-mapping pieces: hundreds of batches)
                saNext = base + 32 + lane < p.Nrep ? packedGet(ix.SA, ix.saBits, p.SAstart + base + 32 + lane) : 0;
                if (lane < nl) {
                    a1 = saCur;
                    aStr = (u32)(a1 >> ix.GstrandBit);
                    a1 &= ix.GstrandMask;
                    aRstart = p.rStart;
                    if (p.Dir == 1 && aStr == 0) { aStr = 1; aRstart = Lread - (aLength + aRstart); }
                    else if (p.Dir == 0 && aStr == 1) { aRstart = Lread - (aLength + aRstart); a1 = ix.nGenome - (aLength + a1); }
                    else if (p.Dir == 1 && aStr == 1) { aStr = 0; a1 = ix.nGenome - (aLength + a1); }
                    kind = 1;
                    if (a1 >= ix.sjGstart) {
                        u64 a1D;
                        if (sjAlignSplit(ix, a1, aLength, a1D, aLengthD, a1A, aLengthA, isj)) { a1 = a1D; kind = 2; } else { kind = 0; isj = SJA_NONE; }
                    }
                    bool lookup = kind != 0;
                    if (kind && useFilter) {
                        const u32 hD = filterHash(aStr, a1 >> P.winBinNbits);
                        lookup = (binFilter[hD >> 5] >> (hD & 31)) & 1u;
                        if (!lookup && kind == 2) { const u32 hA = filterHash(aStr, a1A >> P.winBinNbits); lookup = (binFilter[hA >> 5] >> (hA & 31)) & 1u; }
                    }
                    if (lookup) {   // window lookup: windows do not change during the assignment phase
                        const u64 binD = a1 >> P.winBinNbits, binA = a1A >> P.winBinNbits;
                        bool scan = !sortedLookup;
                        if (sortedLookup) {
                            wD = findWindow(aStr, binD);
                            if (kind == 2) wA = findWindow(aStr, binA);
                            if (wD == -2 || wA == -2) { wD = -1; wA = -1; scan = true; }
                        }
                        #pragma unroll 1
                        for (u32 w = 0; scan && w < ww.nW; w++) {
                            const Window W = swin[w];
                            if (W.Str != aStr) continue;
                            if (wD < 0 && W.gStart <= binD && binD <= W.gEnd) wD = (int)w;
                            if (kind == 2 && wA < 0 && W.gStart <= binA && binA <= W.gEnd) wA = (int)w;
                        }
                    }
                }
                // loci that fall into no window (the majority for multi-mapping pieces) are no-ops of assignAlignToWindow: only the others are visited, in order
                saEnum += nl;
                u32 todo = __ballot_sync(0xffffffffu, kind == 1 ? wD >= 0 : (kind == 2 && (wD >= 0 || wA >= 0)));
                #pragma unroll 1
                while (todo) {
                    const u32 q = (u32)__ffs((int)todo) - 1;
                    todo &= todo - 1;
                    const u32 kq = __shfl_sync(0xffffffffu, kind, q);
                    const u64 a1q = __shfl_sync(0xffffffffu, a1, q);
                    const u64 rq = __shfl_sync(0xffffffffu, aRstart, q);
                    const u32 sq = __shfl_sync(0xffffffffu, aStr, q);
                    const int wDq = __shfl_sync(0xffffffffu, wD, q);
                    if (kq == 1) {
                        if (!coopAssign(ln, ww, wDq, a1q, aLength, aNrep, aFrag, rq, aAnchor, SJA_NONE)) { tooManyAnchors = true; saEnum -= nl - 1 - q; break; }   // (the loci behind q are not visited)
                    } else {
                        const u64 a1Aq = __shfl_sync(0xffffffffu, a1A, q);
                        const u64 lDq = __shfl_sync(0xffffffffu, aLengthD, q), lAq = __shfl_sync(0xffffffffu, aLengthA, q);
                        const u32 isjq = __shfl_sync(0xffffffffu, isj, q);
                        const int wAq = __shfl_sync(0xffffffffu, wA, q);
                        if (!coopAssign(ln, ww, wDq, a1q, lDq, aNrep, aFrag, rq, aAnchor, isjq)) { tooManyAnchors = true; saEnum -= nl - 1 - q; break; }
                        if (!coopAssign(ln, ww, wAq, a1Aq, lAq, aNrep, aFrag, rq + lDq, aAnchor, isjq)) { tooManyAnchors = true; saEnum -= nl - 1 - q; break; }
                    }
                }
            }
        }
        if (tooManyAnchors) ww.nW = 0;
        nWin = ww.nW;
        ln.saEnum = saEnum;
    }
}

// Modes: heavyPool != NULL : the lane of stitch_kernel that owned the read exported its windows + seeds (DFS-heavy read);
//        heavyPool == NULL : the read was routed here right after seeding because it has many loci (nA): the warp also does the
//                            window creation / seed assignment cooperatively (coalesced SA loads, 32-wide scans).
__global__ void __launch_bounds__(128, STITCH_MIN_BLOCKS) stitch_heavy_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                     ReadInfo* __restrict__ info, const Piece* __restrict__ pieces, u32 nHeavy, const u32* __restrict__ heavyList,
                                                     const u64* __restrict__ heavyOff, const u8* __restrict__ heavyPool, u32* __restrict__ counter,
                                                     u8* __restrict__ arenas, Caps caps, star_read_result_t* __restrict__ results,
                                                     star_align_t* __restrict__ staged, u32 smemStride, u8* __restrict__ scratch, HeavyScratch hs) {
    extern __shared__ u8 smem[];
    const u32 lane = threadIdx.x & 31;
    const u32 warpInBlock = threadIdx.x >> 5;
    const u32 warpsPerBlock = blockDim.x >> 5;
    const u32 gwarp = blockIdx.x * warpsPerBlock + warpInBlock;
    // shared memory per warp: R0 | R2 | counters (32 B) | window table
    const u32 perWarp = (2 * smemStride + 32 + caps.maxW * (u32)sizeof(Window) + (caps.maxW + 4) * 4 + ((caps.maxW + 3) & ~3u) + 15) & ~15u;
    u8* R0 = smem + (size_t)warpInBlock * perWarp;
    u8* R2 = R0 + smemStride;
    u32* sh = (u32*)(R0 + 2 * smemStride);   // [0] task ticket, [1] block bump, [2] overflow flag, [3] stored-transcript bump
    Window* swin = (Window*)(R0 + 2 * smemStride + 32);
    u32* taskStart = (u32*)(swin + caps.maxW);            // maxW+1 (+pad): first task of window w (windows without seeds: empty range)
    u8* depthOf = (u8*)(taskStart + caps.maxW + 4);       // maxW
    Lane ln;
    DevTr curL, leafL;
    Frame stackL[STAR_UNDO_DEPTH];
    u8 phL[STAR_DFS_MAX_DEPTH + 4];
    ln.cur = &curL; ln.leaf = &leafL; ln.stack = stackL; ln.ph = phL;
    ln.ix = &ix; ln.P = &P; ln.R0 = R0; ln.R2 = R2; ln.R = R0; ln.caps = caps;
    {   // one arena per WARP: seeds of the windows + the recording state (pool, pointer arrays, compacted window Chr/Str)
        u8* a = arenas + (u64)gwarp * caps.arenaBytes;
        ln.win = (Window*)a; a += (u64)caps.maxW * sizeof(Window);
        ln.wa = (Seed*)a; a += (u64)caps.maxW * caps.spw * sizeof(Seed);
        ln.pool = (DevTr*)a; a += (u64)caps.maxTr * sizeof(DevTr);
        a += 2 * sizeof(DevTr) + (u64)(caps.spw + 2) * sizeof(Frame);
        ln.trPtr = (u16*)a; a += (u64)caps.maxTr * sizeof(u16);
        ln.winBase = (u16*)a; a += (u64)caps.maxW * sizeof(u16);
        ln.winN = (u16*)a;
    }
    u8* ws = scratch + (u64)gwarp * hs.bytesPerWarp;
    const u32 W1 = (hs.maxWin + 2) & ~1u;                                   // even, >= maxWin+1
    TaskOut* taskOut = (TaskOut*)(ws + (u64)W1 * 8 + ((W1 + 7) & ~7u));     // maxTasks (8-byte aligned; the leading area is unused now)
    CandBlock* blocks = (CandBlock*)(taskOut + hs.maxTasks);                // maxBlocks
    u64* trBuf = (u64*)(blocks + hs.maxBlocks);                              // trWords
    u64* epochPtr = trBuf + hs.trWords;                                      // persistent per-warp epoch of the stitch memo
    StitchMemo* memoTab = (StitchMemo*)(epochPtr + 1);                       // memoSlots (zeroed once at allocation)
    const bool memoOn = hs.memoSlots != 0 && caps.maxW <= 4096;
    ln.memo = memoOn ? memoTab : nullptr; ln.memoMask = hs.memoSlots - 1; ln.memoBase = 0; ln.memoHit = 0; ln.memoMiss = 0; ln.lastSeed = -1; ln.coop = 0;
    u64 epoch = *epochPtr;
    WarpWin ww;
    ww.swin = swin; ww.wa = ln.wa; ww.spw = caps.spw; ww.lane = lane; ww.nW = 0;

    long long hc[6] = {0, 0, 0, 0, 0, 0};
    long long eU[3] = {0, 0, 0};
    #pragma unroll 1
    for (;;) {
        long long t0 = clock64();
        u32 k = 0;
        if (lane == 0) k = atomicAdd(counter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= nHeavy) break;
        const u32 i = heavyList[k];
        ReadInfo ri = info[i];
        readBegin(ln, ri);
        const u32 Lread = ri.Lread;
        if (ri.flags || ri.Lread < P.outFilterMatchNmin || ri.Nsplit == 0 || ri.nA == 0) {   // same early exits as stitch_kernel's fetch
            if (lane == 0) {
                if (ri.flags) {
                    star_read_result_t res;
                    res.unmapType = 0; res.nTr = 0; res.nTrOut = 0; res.mapMarker = 0; res.trOffset = 0; res.bestScore = 0; res.bestNMM = 0;
                    res.bestRLength = 0; res.Lread = ri.Lread; res.bestTr = 0;
                    results[i] = res;
                } else if (ri.Lread < P.outFilterMatchNmin) selectExport(ln, ri, i, STAR_MARKER_READ_TOO_SHORT, 0, results, staged, info);
                else if (ri.Nsplit == 0) selectExport(ln, ri, i, STAR_MARKER_NO_GOOD_PIECES, ri.split1_0, results, staged, info);
                else selectExport(ln, ri, i, STAR_MARKER_ALL_PIECES_EXCEED_seedMultimapNmax, ri.multNminL, results, staged, info);
            }
            __syncwarp();
            continue;
        }
        {   // read into shared memory (both orientations)
            const u8* g = reads + (u64)i * stride;
            #pragma unroll 1
            for (u32 b = lane; b < Lread; b += 32) {
                u8 c = g[b];
                R0[b] = c;
                R2[Lread - 1 - b] = c < 4 ? 3 - c : c;
            }
        }
        if (lane == 0) { sh[0] = 0; sh[1] = 0; sh[2] = 0; sh[3] = 0; }
        epoch++;
        if (memoOn && (epoch & 0xFFFFFFULL) == 0) {   // 24-bit epoch wrapped: forget everything
            #pragma unroll 1
            for (u32 q = lane; q < hs.memoSlots; q += 32) memoTab[q].key = 0;
            epoch++;
        }
        __syncwarp();
        u32 nWin = 0;
        u32 overReason = 0;
        warpBuildWindows(ln, ww, ix, P, ri, i, i, pieces, heavyOff, heavyPool, caps, swin, lane, nWin, overReason, taskStart);
        __syncwarp();
        // ---- task table (lane 0): split depth per window, prefix sums (windows without seeds get an empty task range)
        u32 nTasks = 0;
        if (lane == 0 && !overReason) {
            u32 shift = 0;
            #pragma unroll 1
            for (;;) {
                u32 tot = 0;
                #pragma unroll 1
                for (u32 w = 0; w < nWin; w++) {
                    u32 a = swin[w].nWA;
                    u32 d = a <= hs.splitMin ? 0 : (a - hs.splitMin > 8 ? 8 : a - hs.splitMin);
                    d = d > shift ? d - shift : 0;
                    taskStart[w] = tot; depthOf[w] = (u8)d;
                    if (a) tot += 1u << d;
                }
                taskStart[nWin] = tot;
                if (tot <= hs.maxTasks) { nTasks = tot; break; }
                shift++;
            }
        }
        __syncwarp();
        nTasks = __shfl_sync(0xffffffffu, nTasks, 0);
        long long t1 = clock64(); hc[0] += t1 - t0;
        // ---- E phase: lanes evaluate prefix sub-trees, lockstep over {fetch task, DFS node, leaf}
        if (!overReason) {
            u32 ph = 0;   // 0 fetch, 1 node, 2 leaf, 3 idle
            u32 tsk = 0, w = 0, Chr = 0, Str = 0, curBlock = 0xFFFFFFFFu;
            int taskBest = 0;
            const Seed* WA = ln.wa;
            u32 nA = 0;
            #pragma unroll 1
            for (;;) {
                {   // lane-utilisation accounting of the E phase
                    u32 mF = __ballot_sync(0xffffffffu, ph == 0), mN = __ballot_sync(0xffffffffu, ph == 1), mL = __ballot_sync(0xffffffffu, ph == 2);
                    hc[5]++; eU[0] += __popc(mF); eU[1] += __popc(mN); eU[2] += __popc(mL);
                }
                if (ph == 0) {
                    tsk = atomicAdd(&sh[0], 1u);
                    if (tsk >= nTasks || sh[2]) { ph = 3; }
                    else {
                        // window of the task: last w with taskStart[w] <= tsk (skipping empty ranges is automatic: taskStart[w+1] > tsk)
                        u32 lo = 0, hi = nWin;
                        #pragma unroll 1
                        while (lo + 1 < hi) { u32 mid = (lo + hi) >> 1; if (taskStart[mid] <= tsk) lo = mid; else hi = mid; }
                        w = lo;
                        const Window W = swin[w];
                        Chr = W.Chr; Str = W.Str; nA = W.nWA;
                        WA = ln.wa + (u64)w * caps.spw;
                        ln.R = Str == 0 ? R0 : R2;
                        dfsInit(ln);
                        ln.forceDepth = depthOf[w];
                        ln.forceBits = tsk - taskStart[w];
                        ln.memoBase = ((epoch & 0xFFFFFFULL) << 40) | ((u64)w << 28);
                        ln.memo = (memoOn && nA >= 10) ? memoTab : nullptr;   // repetition only pays in windows with many seeds
                        taskOut[tsk].first = 0xFFFFFFFFu; taskOut[tsk].last = 0xFFFFFFFFu;
                        curBlock = 0xFFFFFFFFu;
                        taskBest = 0;
                        ph = 1;
                    }
                }
                if (ph == 1) {
                    int r = dfsStep(ln, WA, nA);
                    if (r == DFS_LEAF) ph = 2;
                    else if (r == DFS_DONE) ph = 0;
                }
                if (ph == 2) {
                    ln.leaves++;
                    if (evalLeaf(ln, ln.leafScore, ln.leafR2, ln.leafG2, Chr, Str, Str)) {
                        if (curBlock == 0xFFFFFFFFu || blocks[curBlock].count == CAND_PER_BLOCK) {
                            u32 nb = atomicAdd(&sh[1], 1u);
                            if (nb >= hs.maxBlocks) { sh[2] = 1; nb = 0xFFFFFFFFu; }
                            else {
                                blocks[nb].next = 0xFFFFFFFFu; blocks[nb].count = 0;
                                if (curBlock == 0xFFFFFFFFu) taskOut[tsk].first = nb; else blocks[curBlock].next = nb;
                                taskOut[tsk].last = nb;
                            }
                            curBlock = nb;
                        }
                        if (curBlock != 0xFFFFFFFFu) {
                            const int sc = ln.leaf->h.maxScore;
                            Cand c; c.mask = ln.inclMask; c.score = (short)sc; c.iFrag = ln.leaf->h.iFrag; c.pad = 0; c.trOff = 0xFFFFFFFFu;
                            if (sc + P.outFilterMultimapScoreRange >= taskBest) {   // likely to be recorded: keep the evaluated transcript
                                if (sc > taskBest) taskBest = sc;
                                const u32 nEx = ln.leaf->h.nExons;
                                const u32 words = (u32)(sizeof(TrHead) / 8) + nEx * (u32)(sizeof(Exon) / 8);
                                u32 off = atomicAdd(&sh[3], words);
                                if (off + words <= hs.trWords) {
                                    u64* dst = trBuf + off;
                                    const u64* sh8 = (const u64*)&ln.leaf->h;
                                    #pragma unroll 1
                                    for (u32 q = 0; q < sizeof(TrHead) / 8; q++) dst[q] = sh8[q];
                                    const u64* se = (const u64*)ln.leaf->ex;
                                    #pragma unroll 1
                                    for (u32 q = 0; q < nEx * (sizeof(Exon) / 8); q++) dst[sizeof(TrHead) / 8 + q] = se[q];
                                    c.trOff = off;
                                }
                            }
                            CandBlock& B = blocks[curBlock];
                            B.c[B.count] = c;
                            B.count++;
                        }
                    }
                    ph = sh[2] ? 3 : 1;
                }
                if (__all_sync(0xffffffffu, ph == 3)) break;
            }
        }
        __syncwarp();
        {   // work counters of the E phase (nodes, leaves) summed over the lanes
            u64 nd = ln.nodes, lv = ln.leaves;
            #pragma unroll 1
            for (int o = 16; o > 0; o >>= 1) { nd += __shfl_down_sync(0xffffffffu, nd, o); lv += __shfl_down_sync(0xffffffffu, lv, o); }
            ln.nodes = nd; ln.leaves = lv;   // meaningful on lane 0
        }
        if (sh[2] != 0 && !overReason) overReason = 5;
        long long t2 = clock64(); hc[1] += t2 - t1; hc[3] += nTasks;
        // ---- R phase: lane 0 replays the order-dependent recording
        if (lane == 0) {
            ln.forceDepth = 0; ln.forceBits = 0;
            if (overReason) {
                ln.overflow = overReason;
            } else {
                #pragma unroll 1
                for (u32 q = 0; q < caps.maxTr; q++) ln.trPtr[q] = (u16)q;
                #pragma unroll 1
                for (u32 w = 0; w < nWin && !ln.overflow; w++) {
                    const Window W = swin[w];
                    if (W.nWA == 0) continue;
                    u16* wTr = nullptr; u16 nWinTr = 0;
                    int rc = windowBegin(ln, wTr, nWinTr);
                    if (rc == 2) { ln.overflow = 3; break; }
                    if (rc == 1) break;
                    const u32 Chr = W.Chr, Str = W.Str, nA = W.nWA;
                    const Seed* WA = ln.wa + (u64)w * caps.spw;
                    ln.R = Str == 0 ? R0 : R2;
                    ln.memoBase = ((epoch & 0xFFFFFFULL) << 40) | ((u64)w << 28);
                    ln.memo = (memoOn && nA >= 10) ? memoTab : nullptr;
                    #pragma unroll 1
                    for (u32 t = taskStart[w]; t < taskStart[w + 1] && !ln.overflow; t++) {
                        u32 b = taskOut[t].first;
                        #pragma unroll 1
                        while (b != 0xFFFFFFFFu && !ln.overflow) {
                            const CandBlock& B = blocks[b];
                            #pragma unroll 1
                            for (u32 q = 0; q < B.count; q++) {
                                const Cand c = B.c[q];
                                if (c.iFrag >= 0 && ln.maxScoreMate[c.iFrag] < c.score) ln.maxScoreMate[c.iFrag] = c.score;
                                int wBest = ln.pool[wTr[0]].h.maxScore;
                                if (c.score + P.outFilterMultimapScoreRange >= wBest ||
                                    (c.iFrag >= 0 && c.score + P.outFilterMultimapScoreRange >= ln.maxScoreMate[c.iFrag])) {
                                    if (nWinTr > caps.maxTr - ln.trNtotal - 1) { ln.overflow = 3; break; }
                                    if (c.trOff != 0xFFFFFFFFu) {   // transcript stored by the lane that evaluated the leaf
                                        const u64* src = trBuf + c.trOff;
                                        u64* dh = (u64*)&ln.leaf->h;
                                        #pragma unroll 1
                                        for (u32 z = 0; z < sizeof(TrHead) / 8; z++) dh[z] = src[z];
                                        const u32 nEx = ln.leaf->h.nExons;
                                        u64* de = (u64*)ln.leaf->ex;
                                        #pragma unroll 1
                                        for (u32 z = 0; z < nEx * (sizeof(Exon) / 8); z++) de[z] = src[sizeof(TrHead) / 8 + z];
                                        recordLeaf(ln, wTr, &nWinTr);
                                    } else {
                                        int Score; u32 tR2; u64 tG2;
                                        hc[4]++;
                                        bool ok = replayPath(ln, WA, nA, c.mask, Score, tR2, tG2) && evalLeaf(ln, Score, tR2, tG2, Chr, Str, Str);
                                        if (ok) recordLeaf(ln, wTr, &nWinTr);
                                    }
                                }
                            }
                            b = B.next;
                        }
                    }
                    // windowEnd compacts Chr/Str into ln.win[] (global arena), independent of the shared-memory table
                    windowEnd(ln, Chr, Str, wTr, nWinTr);
                }
            }
            selectExport(ln, ri, i, 0, 0, results, staged, info);
        }
        __syncwarp();
        hc[2] += clock64() - t2;
    }
    if (lane == 0) *epochPtr = epoch;
    #pragma unroll 1
    for (int q = 0; q < 6; q++) PROF_ADD(16 + q, hc[q]);
    #pragma unroll 1
    for (int q = 0; q < 3; q++) PROF_ADD(22 + q, eU[q]);
    {
        u64 hsum = ln.memoHit, msum = ln.memoMiss;
        #pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) { hsum += __shfl_down_sync(0xffffffffu, hsum, o); msum += __shfl_down_sync(0xffffffffu, msum, o); }
        PROF_ADD(25, hsum); PROF_ADD(26, msum);
    }
}

#include "stitch_flat.cuh"

__global__ void prof_read_kernel(unsigned long long* out, int reset) {
    int t = threadIdx.x;
    if (t < 32) { out[t] = g_prof[t]; if (reset) g_prof[t] = 0; }
}

// Heaviest-first schedule of the fast path: key = ~nA (number of genomic loci of the stored pieces, known after seeding).
// The per-read work is heavy-tailed (a read inside a repeat family costs 1000x the median); starting those reads first keeps
// the tail of the persistent kernel short.
__global__ void order_keys_kernel(const ReadInfo* __restrict__ info, u32 nReads, u32* __restrict__ keys, u32* __restrict__ vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nReads) { keys[i] = 0xFFFFFFFFu - info[i].nA; vals[i] = i; }
}

// number of reads routed to the warp-per-read kernel right after seeding (many genomic loci)
__global__ void count_heavy_kernel(const ReadInfo* __restrict__ info, u32 nReads, u32 naLimit, u32* __restrict__ count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    bool h = i < nReads && info[i].nA > naLimit;
    u32 m = __ballot_sync(0xffffffffu, h);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// Compaction: results[i].trOffset = exclusive prefix sum of nTrOut; aligns[trOffset+k] = staged[i*nOut+k].
// One warp per read copies its alignments with coalesced 16-byte words.
__global__ void pack_kernel(const star_read_result_t* __restrict__ results, const u64* __restrict__ offsets, const star_align_t* __restrict__ staged,
                            u32 nOut, u32 nReads, star_align_t* __restrict__ aligns) {
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    const u32 words = sizeof(star_align_t) / 16;
    #pragma unroll 1
    for (u32 i = warp; i < nReads; i += nWarps) {
        u32 n = results[i].nTrOut;
        if (n == 0) continue;
        const uint4* s = (const uint4*)(staged + (u64)i * nOut);
        uint4* d = (uint4*)(aligns + offsets[i]);
        #pragma unroll 1
        for (u32 w = lane; w < n * words; w += 32) d[w] = s[w];
    }
}

// exclusive scan of nTrOut over the chunk (trOffset of every read, total number of records): per-block sums, a scan of the block sums,
// then every block writes the offsets of its contiguous range of reads (tiles of blockDim reads, scanned with warp shuffles)
__global__ void __launch_bounds__(256) scan_partial_kernel(const star_read_result_t* __restrict__ results, u32 nReads, u32 per, u64* __restrict__ partial) {
    __shared__ u64 ws[8];
    const u32 lo = blockIdx.x * per, hi = lo + per < nReads ? lo + per : nReads;
    u64 s = 0;
    for (u32 i = lo + threadIdx.x; i < hi; i += blockDim.x) s += results[i].nTrOut;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (u32 w = 0; w < (blockDim.x >> 5); w++) t += ws[w]; partial[blockIdx.x] = t; }
}
__global__ void scan_top_kernel(u64* __restrict__ partial, u32 nBlocks, u64* __restrict__ total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        u64 run = 0;
        #pragma unroll 1
        for (u32 k = 0; k < nBlocks; k++) { const u64 v = partial[k]; partial[k] = run; run += v; }
        *total = run;
    }
}
__global__ void __launch_bounds__(256) scan_write_kernel(star_read_result_t* __restrict__ results, u64* __restrict__ offsets, u32 nReads, u32 per, const u64* __restrict__ partial) {
    __shared__ u64 ws[8];
    __shared__ u64 carry;
    const u32 lo = blockIdx.x * per, hi = lo + per < nReads ? lo + per : nReads;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = partial[blockIdx.x];
    __syncthreads();
    #pragma unroll 1
    for (u32 base = lo; base < hi; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        const u64 v = i < hi ? results[i].nTrOut : 0;
        u64 incl = v;
        for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += x; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        u64 before = carry;
        for (u32 w = 0; w < warp; w++) before += ws[w];
        if (i < hi) { const u64 off = before + incl - v; offsets[i] = off; results[i].trOffset = off; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + incl;
        __syncthreads();
    }
}

}  // namespace starb
