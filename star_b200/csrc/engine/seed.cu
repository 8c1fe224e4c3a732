// seed.cu — read preparation + maximal-mappable-prefix seed search kernels (sm_100a).
//
// What is computed (bit-exact with the reference, proven by tests/ against oracle/):
//   prep_reads_kernel   ReadAlign::oneRead prologue          reference source/ReadAlign_oneRead.cpp:35-78,
//                       convertNucleotidesToNumbers / complementSeqNumbers  SequenceFuns.cpp:4-14,131-146
//   seed_search_kernel  qualitySplit                          SequenceFuns.cpp:411-444
//                       ReadAlign::mapOneRead search schedule ReadAlign_mapOneRead.cpp:37-93
//                       maxMappableLength2strands             ReadAlign_maxMappableLength2strands.cpp:5-115
//                       maxMappableLength / findMultRange / compareSeqToGenome   SuffixArrayFuns.cpp:4-207
//                       storeAligns                           ReadAlign_storeAligns.cpp:10-160
// How (B200-first, not the reference's structure): one persistent lane per read; the combined read is staged
// once in shared memory (conflict-free odd-word stride), the packed SA / SAi words are fetched as aligned
// 64-bit read-only loads, and the pieces of a read are kept insertion-sorted in a 16-byte-per-piece
// global slab.  The kernel is latency/HBM-sector bound (random 8-byte SA probes + short genome runs);
// DESIGN.md states its algorithmic bytes and roofline.
#include "dev.cuh"
#include "seed_warp.cuh"

namespace starb {

__device__ __forceinline__ u8 nt2num(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// One warp per read: writes Read1[0] = mate1 | 11 | revcomp(mate2) into reads[i*stride ..].
__global__ void prep_reads_kernel(const char* __restrict__ seq, const u64* __restrict__ seqOff, u32 nReads, u32 nMates,
                                  u8* __restrict__ reads, u32 stride, ReadInfo* __restrict__ info, star_params_t P) {
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    u32 lane = threadIdx.x & 31;
    u32 nWarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = warp; i < nReads; i += nWarps) {
        const u64* off = seqOff + (u64)i * nMates;
        u64 o0 = off[0], o1 = off[1];
        u32 l0 = (u32)(o1 - o0);
        u32 l1 = nMates == 2 ? (u32)(off[2] - o1) : 0;
        u8* r = reads + (u64)i * stride;
        for (u32 k = lane; k < l0; k += 32) r[k] = nt2num(seq[o0 + k]);
        u32 Lread = l0;
        if (nMates == 2) {
            if (lane == 0) r[l0] = STAR_MARK_FRAG_SPACER_BASE;
            for (u32 k = lane; k < l1; k += 32) {
                u8 c = nt2num(seq[o1 + l1 - 1 - k]);
                r[l0 + 1 + k] = c < 4 ? 3 - c : c;
            }
            Lread = l0 + l1 + 1;
        }
        if (lane == 0) {
            ReadInfo ri;
            ri.Lread = Lread;
            ri.readLength[0] = (u16)l0; ri.readLength[1] = (u16)l1;
            ri.nP = 0; ri.nA = 0; ri.mapMarker = 0; ri.multNminL = 0; ri.Nsplit = 0; ri.split1_0 = 0; ri.flags = 0;
            u64 a = P.outFilterMismatchNmax;
            u64 b = (u64)(P.outFilterMismatchNoverReadLmax * (double)(l0 + l1));   // ReadAlign_oneRead.cpp:78
            ri.outFilterMismatchNmaxTotal = (u32)(a < b ? a : b);
            info[i] = ri;
        }
    }
}

struct SeedCtx {
    const DevIndex* ix;
    const u8* R;      // Read1[0] of this read (shared memory)
    u64 searches, saiWords, compareCalls, basesExamined;
};

// SuffixArrayFuns.cpp:10-104
__device__ __forceinline__ u64 compareSeqToGenome(SeedCtx& c, u64 S, u64 N, u64 L, u64 iSA, bool dirR, bool& compRes) {
    const DevIndex& ix = *c.ix;
    u64 SAstr = packedGet(ix.SA, ix.saBits, iSA);
    bool dirG = (SAstr >> ix.GstrandBit) == 0;
    SAstr &= ix.GstrandMask;
    c.compareCalls++;
    const u8* G = ix.G;
    u64 n = N - L;
    if (dirG) {
        const u8* g = G + SAstr + L;
        if (dirR) {
            const u8* s = c.R + S + L;
            for (u64 ii = 0; ii < n; ii++) {
                u8 sv = s[ii], gv = __ldg(g + ii);
                if (sv != gv) { compRes = sv > gv; c.basesExamined += ii + 1; return ii + L; }
            }
        } else {
            const u8* s = c.R + S - L;
            for (u64 ii = 0; ii < n; ii++) {
                u8 sv = 3 - *(s - ii), gv = __ldg(g + ii);   // Read1[1] = complement; piece bases are always 0..3
                if (sv != gv) { compRes = sv > gv; c.basesExamined += ii + 1; return ii + L; }
            }
        }
    } else {
        const u8* g = G + (ix.nGenome - 1 - SAstr - L);
        if (dirR) {
            const u8* s = c.R + S + L;
            for (u64 ii = 0; ii < n; ii++) {
                u8 sv = 3 - s[ii], gv = __ldg(g - ii);
                if (sv != gv) { compRes = !(sv > gv || gv > 3); c.basesExamined += ii + 1; return ii + L; }
            }
        } else {
            const u8* s = c.R + S - L;
            for (u64 ii = 0; ii < n; ii++) {
                u8 sv = *(s - ii), gv = __ldg(g - ii);
                if (sv != gv) { compRes = !(sv > gv || gv > 3); c.basesExamined += ii + 1; return ii + L; }
            }
        }
    }
    c.basesExamined += n;
    return N;
}

__device__ __forceinline__ u64 medianUint2(u64 a, u64 b) { return a / 2 + b / 2 + (a % 2 + b % 2) / 2; }

// SuffixArrayFuns.cpp:106-131
__device__ u64 findMultRange(SeedCtx& c, u64 i3, u64 L3, u64 i1, u64 L1, u64 i1a, u64 L1a, u64 i1b, u64 L1b, bool dirR, u64 S) {
    bool compRes;
    if (L1 < L3) {
        L1b = L1; i1b = i1; i1a = i3;
    } else {
        if (L1a < L1) { L1b = L1a; i1b = i1a; i1a = i1; }
    }
    while ((i1b + 1 < i1a) | (i1b > i1a + 1)) {
        u64 i1c = medianUint2(i1a, i1b);
        u64 L1c = compareSeqToGenome(c, S, L3, L1b, i1c, dirR, compRes);
        if (L1c == L3) i1a = i1c;
        else { i1b = i1c; L1b = L1c; }
    }
    return i1a;
}

// SuffixArrayFuns.cpp:133-207
__device__ u64 maxMappableLength(SeedCtx& c, u64 S, u64 N, u64 i1, u64 i2, bool dirR, u64& L, u64* indStartEnd) {
    bool compRes = false;
    u64 L1, L2, i3, L3, L1a, L1b, L2a, L2b, i1a, i1b, i2a, i2b;
    L1 = compareSeqToGenome(c, S, N, L, i1, dirR, compRes);
    L2 = compareSeqToGenome(c, S, N, L, i2, dirR, compRes);
    L = L1 < L2 ? L1 : L2;
    L1a = L1; L1b = L1; i1a = i1; i1b = i1;
    L2a = L2; L2b = L2; i2a = i2; i2b = i2;
    i3 = i1; L3 = L1;
    while (i1 + 1 < i2) {
        i3 = medianUint2(i1, i2);
        L3 = compareSeqToGenome(c, S, N, L, i3, dirR, compRes);
        if (L3 == N) break;
        if (compRes) {
            if (L3 > L1) { L1b = L1a; L1a = L1; i1b = i1a; i1a = i1; }
            i1 = i3; L1 = L3;
        } else {
            if (L3 > L2) { L2b = L2a; L2a = L2; i2b = i2a; i2a = i2; }
            i2 = i3; L2 = L3;
        }
        L = L1 < L2 ? L1 : L2;
    }
    if (L3 < N) {
        if (L1 > L2) { i3 = i1; L3 = L1; } else { i3 = i2; L3 = L2; }
    }
    i1 = findMultRange(c, i3, L3, i1, L1, i1a, L1a, i1b, L1b, dirR, S);
    i2 = findMultRange(c, i3, L3, i2, L2, i2a, L2a, i2b, L2b, dirR, S);
    L = L3;
    indStartEnd[0] = i1; indStartEnd[1] = i2;
    return i2 - i1 + 1;
}

struct StoreState {
    Piece* PC;
    u32 nP, maxP;
    u32 nA;
    u32 multNmin, multNminL;
    u32 flags;
};

// ReadAlign_storeAligns.cpp:10-51,140-160 (OPTIM_STOREaligns_SIMPLE)
__device__ void storeAligns(StoreState& st, const star_params_t& P, u32 iDir, u64 Shift, u64 Nrep, u64 L, u64 SAstart, u32 iFrag) {
    if (Nrep > P.seedMultimapNmax) {
        if (Nrep < st.multNmin || st.multNmin == 0) { st.multNmin = (u32)Nrep; st.multNminL = (u32)L; }
        return;
    }
    st.nA += (u32)Nrep;
    u32 rStart = (u32)(iDir == 0 ? Shift : Shift + 1 - L);
    int iP;
    for (iP = (int)st.nP - 1; iP >= 0; iP--) {
        u32 r0 = st.PC[iP].rStart;
        if (r0 <= rStart) {
            if (r0 == rStart && st.PC[iP].Length < L) continue;
            if (r0 == rStart && st.PC[iP].Length == L) return;
            break;
        }
    }
    iP = iP + 1;
    if (st.nP + 1 > P.seedPerReadNmax) { st.flags |= 2; return; }   // fatal in the reference (:46-51)
    if (st.nP + 1 > st.maxP) { st.flags |= 1; return; }             // fast-path slab full: redo on the slow path
    for (int ii = (int)st.nP - 1; ii >= iP; ii--) st.PC[ii + 1] = st.PC[ii];
    st.nP++;
    Piece p;
    p.SAstart = SAstart; p.rStart = (u16)rStart; p.Length = (u16)L; p.Nrep = (u16)Nrep; p.Dir = (u8)iDir; p.iFrag = (u8)iFrag;
    st.PC[iP] = p;
    if (Nrep != 1) {
        if (Nrep < st.multNmin || st.multNmin == 0) { st.multNmin = (u32)Nrep; st.multNminL = (u32)L; }
    }
}

// ReadAlign_maxMappableLength2strands.cpp:5-115 with gSAsparseD == 1
__device__ void maxMappableLength2strands(SeedCtx& c, StoreState& st, const star_params_t& P, u64 pieceStart, u64 pieceLength, u32 iDir,
                                          u64& maxLbest, u32 iFrag) {
    const DevIndex& ix = *c.ix;
    u64 Nrep = 0, indStartEnd[2] = {0, 0}, maxL = 0;
    bool dirR = iDir == 0;
    c.searches++;
    u64 Lmax = ix.gSAindexNbases < pieceLength ? ix.gSAindexNbases : pieceLength;
    u64 ind1 = 0;
    if (dirR) {
        for (u64 ii = 0; ii < Lmax; ii++) { ind1 <<= 2; ind1 += (u64)c.R[pieceStart + ii]; }
    } else {
        for (u64 ii = 0; ii < Lmax; ii++) { ind1 <<= 2; ind1 += (3 - (u64)c.R[pieceStart - ii]); }
    }
    u64 Lind = Lmax;
    u64 iSA1 = 0, iSA2 = 0;
    while (Lind > 0) {
        iSA1 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1);
        c.saiWords++;
        if ((iSA1 & ix.SAiMarkAbsentMaskC) == 0) break;
        --Lind;
        ind1 = ind1 >> 2;
    }
    bool iSA2good = true;
    if (ix.genomeSAindexStart[Lind - 1] + ind1 + 1 < ix.genomeSAindexStart[Lind]) {
        iSA2 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1 + 1);
        c.saiWords++;
        if ((iSA2 & ix.SAiMarkAbsentMaskC) == 0) {
            iSA2 = (iSA2 & ix.SAiMarkNmask) - 1;
        } else {
            iSA2 = ix.nSA - 1;
            iSA2good = false;
        }
    } else {
        iSA2 = ix.nSA - 1;
        iSA2good = false;
    }
    bool iSA1noN = (iSA1 & ix.SAiMarkNmaskC) == 0;
    if (Lind < ix.gSAindexNbases && iSA1noN && iSA2good) {
        indStartEnd[0] = iSA1; indStartEnd[1] = iSA2;
        Nrep = iSA2 - iSA1 + 1;
        maxL = Lind;
    } else if (iSA1 == iSA2 && iSA1noN && iSA2good) {
        indStartEnd[0] = indStartEnd[1] = iSA1;
        Nrep = 1;
        bool cr;
        maxL = compareSeqToGenome(c, pieceStart, pieceLength, Lind, iSA1, dirR, cr);
    } else {
        maxL = (iSA2good && iSA1noN) ? Lind : 0;
        Nrep = maxMappableLength(c, pieceStart, pieceLength, iSA1 & ix.SAiMarkNmask, iSA2, dirR, maxL, indStartEnd);
    }
    maxLbest = maxL;
    storeAligns(st, P, iDir, pieceStart, Nrep, maxL, indStartEnd[0], iFrag);
}

// Persistent lanes: each lane takes the next read from `counter` (or from readList on the slow path).
__global__ void __launch_bounds__(128) seed_search_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                          ReadInfo* __restrict__ info, Piece* __restrict__ pieces, u32 maxP, u32 nReads,
                                                          const u32* __restrict__ readList, u32* __restrict__ counter,
                                                          WorkCounters* __restrict__ /*wc*/, u32 smemStride) {
    extern __shared__ u8 smem[];
    u8* R = smem + (size_t)threadIdx.x * smemStride;
    SeedCtx c;
    c.ix = &ix; c.R = R;
    for (;;) {
        u32 k = atomicAdd(counter, 1u);
        if (k >= nReads) break;
        c.searches = 0; c.saiWords = 0; c.compareCalls = 0; c.basesExamined = 0;
        u32 i = readList ? readList[k] : k;
        ReadInfo ri = info[i];
        u32 Lread = ri.Lread;
        const u8* src = reads + (u64)i * stride;
        for (u32 b = 0; b < Lread; b++) R[b] = src[b];
        // qualitySplit SequenceFuns.cpp:411-444
        u32 splitStart[10], splitLen[10], splitFrag[10];
        u32 Nsplit = 0;
        {
            u32 iR = 0, iS = 0, LgoodMin = 0, iFrag = 0;
            u32 maxNsplit = (u32)(P.maxNsplit < 10 ? P.maxNsplit : 10);
            while ((iR < Lread) & (iS < maxNsplit)) {
                while (iR < Lread && R[iR] > 3) {
                    if (R[iR] == STAR_MARK_FRAG_SPACER_BASE) iFrag++;
                    iR++;
                }
                if (iR == Lread) break;
                u32 iR1 = iR;
                while (iR < Lread && R[iR] <= 3) iR++;
                if ((iR - iR1) > LgoodMin) LgoodMin = iR - iR1;
                if ((iR - iR1) < P.seedSplitMin) continue;
                splitStart[iS] = iR1; splitLen[iS] = iR - iR1; splitFrag[iS] = iFrag;
                iS++;
            }
            Nsplit = iS;
            ri.Nsplit = (u16)iS;
            ri.split1_0 = (u16)(iS == 0 ? LgoodMin : splitLen[0]);
        }
        StoreState st;
        st.PC = pieces + (u64)k * maxP;   // slab index = position in this launch (k), not the read id
        st.nP = 0; st.maxP = maxP; st.nA = 0; st.multNmin = 0; st.multNminL = 0; st.flags = 0;
        // ReadAlign_mapOneRead.cpp:37-93
        u64 a = P.seedSearchStartLmax;
        u64 b = (u64)(P.seedSearchStartLmaxOverLread * (double)(Lread - 1));
        u64 seedSearchStartLmax = a < b ? a : b;
        for (u32 ip = 0; ip < Nsplit && !st.flags; ip++) {
            u64 pl = splitLen[ip], ps = splitStart[ip];
            u64 Nstart = (P.seedSearchStartLmax > 0 && seedSearchStartLmax < pl) ? pl / seedSearchStartLmax + 1 : 1;
            u64 Lstart = pl / Nstart;
            bool flagDirMap = true;
            for (u32 iDir = 0; iDir < 2; iDir++) {
                for (u64 istart = 0; istart < Nstart; istart++) {
                    if (flagDirMap || istart > 0) {
                        u64 Lmapped = 0;
                        while (istart * Lstart + Lmapped + P.seedMapMin < pl) {
                            u64 Shift = iDir == 0 ? (ps + istart * Lstart + Lmapped) : (ps + pl - istart * Lstart - 1 - Lmapped);
                            u64 seedLength = pl - Lmapped - istart * Lstart;
                            u64 L;
                            maxMappableLength2strands(c, st, P, Shift, seedLength, iDir, L, splitFrag[ip]);
                            if (iDir == 0 && istart == 0 && Lmapped == 0 && Shift + L == pl) flagDirMap = false;
                            Lmapped += L;
                            if (st.flags) break;
                        }
                    }
                    if (__builtin_expect(P.seedSearchLmax > 0, 0) && !st.flags) {   // ReadAlign_mapOneRead.cpp:81-87: fixed-length search from every start (off by default)
                        const u64 Shift = iDir == 0 ? (ps + istart * Lstart) : (ps + pl - istart * Lstart - 1);
                        const u64 room = iDir == 0 ? (ps + pl - Shift) : (Shift + 1);
                        u64 L;
                        maxMappableLength2strands(c, st, P, Shift, P.seedSearchLmax < room ? P.seedSearchLmax : room, iDir, L, splitFrag[ip]);
                    }
                    if (st.flags) break;
                }
                if (st.flags) break;
            }
        }
        ri.nP = (u16)st.nP;
        ri.nA = st.nA;
        ri.multNminL = st.multNminL;
        ri.flags = st.flags;
        ri.cSearches = (u32)c.searches; ri.cSaiWords = (u32)c.saiWords; ri.cCompare = (u32)c.compareCalls; ri.cBases = (u32)c.basesExamined;
        ri.cSaEnum = 0; ri.cNodes = 0; ri.cLeaves = 0; ri.cSlow = readList ? 1 : 0;
        info[i] = ri;
    }
}

// Warp-uniform variant (seed_warp.cuh): one read per WARP, the search of a start interval narrows with 32 probes per step and
// examines windows of <= 32 SA rows in one step.  Same stored pieces as seed_search_kernel (checked lane by lane on the CPU by
// tests/test_warp_emulation.py).  Opt-in (STAR_B200_SEED_WARP=1) until it has been measured on the GPU.  Work counters of this
// kernel: searches and SAindex words as the reference; cCompare = SA rows probed, cBases = genome bases examined by all lanes.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) seed_search_warp_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P,
                                                                    const u8* __restrict__ reads, u32 stride, ReadInfo* __restrict__ info,
                                                                    Piece* __restrict__ pieces, u32 maxP, u32 nReads, const u32* __restrict__ readList,
                                                                    u32* __restrict__ counter, u32 smemStride) {
    extern __shared__ u8 smem[];
    u8* R = smem + 16 + (size_t)(threadIdx.x >> 5) * smemStride;   // 16 bytes of slack before the first and after the last row (8-byte gathers)
    const DevWarp w;
    #pragma unroll 1
    for (;;) {
        u32 k = 0;
        if (w.lane == 0) k = atomicAdd(counter, 1u);
        k = w.shfl(k, 0);
        if (k >= nReads) break;
        const u32 i = readList ? readList[k] : k;
        ReadInfo ri = info[i];
        const u32 Lread = ri.Lread;
        const u8* src = reads + (u64)i * stride;
        w.sync();
        #pragma unroll 1
        for (u32 b = w.lane; b < Lread; b += 32) R[b] = src[b];
        w.sync();
        SeedWarpOut st;
        st.PC = pieces + (u64)k * maxP;   // slab index = position in this launch (k), not the read id
        st.maxP = maxP;
        warpSeedRead<DevWarp>(w, ix, P, R, Lread, st);
        const u32 bases = w.reduceAdd(st.basesLane);
        if (w.lane == 0) {
            ri.Nsplit = (u16)st.Nsplit; ri.split1_0 = (u16)st.split1_0;
            ri.nP = (u16)st.nP; ri.nA = st.nA; ri.multNminL = st.multNminL; ri.flags = st.flags;
            ri.cSearches = st.searches; ri.cSaiWords = st.saiWords; ri.cCompare = st.probes; ri.cBases = bases;
            ri.cSaEnum = 0; ri.cNodes = 0; ri.cLeaves = 0; ri.cSlow = readList ? 1 : 0;
            info[i] = ri;
        }
    }
}

#ifndef STAR_CUDA_HOST_SHIM   // (kernel launches need nvcc; the host emulation of the tests calls the kernels directly)
void launch_seed_warp(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info,
                      Piece* pieces, u32 maxP, u32 nReads, const u32* readList, u32* counter, u32 smemStride) {
    const u32 smem = 4 * smemStride + 32;
    if (ctasPerSM <= 6) seed_search_warp_kernel<6><<<nSM * 6, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, maxP, nReads, readList, counter, smemStride);
    else if (ctasPerSM <= 8) seed_search_warp_kernel<8><<<nSM * 8, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, maxP, nReads, readList, counter, smemStride);
    else seed_search_warp_kernel<12><<<nSM * 12, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, maxP, nReads, readList, counter, smemStride);
}
#endif

// sums the per-read work counters (one warp-reduced atomic per counter per warp)
__global__ void reduce_counters_kernel(const ReadInfo* __restrict__ info, u32 nReads, WorkCounters* __restrict__ wc) {
    u64 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nReads; i += gridDim.x * blockDim.x) {
        const ReadInfo ri = info[i];
        v[0] += ri.cSearches; v[1] += ri.cSaiWords; v[2] += ri.cCompare; v[3] += ri.cBases; v[4] += ri.cSaEnum; v[5] += ri.cNodes; v[6] += ri.cLeaves; v[7] += ri.cSlow;
    }
    u64* out = (u64*)wc;
    for (int k = 0; k < 8; k++) {
        u64 x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, x);
    }
}

}  // namespace starb
