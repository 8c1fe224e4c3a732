// sjdb_kernels.cuh — device part of the on-the-fly junction insertion (reference source/sjdbBuildIndex.cpp).
//
//   sjdb_search_kernel   one thread per suffix of the junction inserts (both strands): the SA row in front of which the suffix has to
//                        be inserted = suffixArraySearch1 (SuffixArrayFuns.cpp:309-351) with compareSeqToGenome1 (:233-306): a binary
//                        search over the whole SA in which a spacer (code 5) of the insert sorts behind the genome's (compareRefEnds,
//                        :221-231, with gInsert = -1: new rows always go behind equal old ones).
//   sjdb_merge_sa_kernel one thread per 64 rows of the NEW suffix array (= GstrandBit+1 whole 64-bit words, so no two threads share a
//                        word), a warp's 32 groups staged in shared memory and stored coalesced: old rows are re-based (reverse-strand coordinates grow with the genome, rows of old inserts move with
//                        their junction's new index) and the sorted new rows are spliced in (sjdbBuildIndex.cpp:141-214).
//                        HBM-bound: reads nSA x (GstrandBit+1)/8 bytes, writes as much plus the new rows.
// Both are grid-stride loops, so that the host emulation (oracle/engine_emul.cpp) can run them as one CTA.
#pragma once
#include "dev.cuh"

namespace starb {

struct SjdbIndex {
    const u8* G;       // base 0 of the old genome, 256 bytes of code 5 readable on both sides
    const u64* SA;     // packed words of the old suffix array
    u64 nGenome, nSA;
    u32 GstrandBit, saBits;
};

// compareSeqToGenome1 from offset L on: +1 / -1 = the insert suffix s sorts behind / in front of SA row iSA; Lout = common length
__device__ __forceinline__ int sjdbCompare(const SjdbIndex& ix, const u8* __restrict__ s, u64 L, u64 iSA, u64& Lout) {
    u64 SAstr = packedGet(ix.SA, ix.saBits, iSA);
    const bool dirG = (SAstr >> ix.GstrandBit) == 0;
    SAstr &= ~(1ULL << ix.GstrandBit);
    if (dirG) {
        const u8* g = ix.G + SAstr;
#pragma unroll 1
        for (u64 ii = L;; ii++) {
            const u8 a = SB_LDG(s + ii), b = SB_LDG(g + ii);
            if (a != b) { Lout = ii; return a > b ? 1 : -1; }
            if (a == 5) { Lout = ii; return 1; }     // both at a spacer: the insert is the later text
        }
    } else {   // row on the reverse strand: the genome is read backwards and complemented
        const u8* g = ix.G + (ix.nGenome - 1 - SAstr);
#pragma unroll 1
        for (u64 ii = L;; ii++) {
            const u8 a = SB_LDG(s + ii);
            u8 b = SB_LDG(g - (i64)ii);
            if (b < 4) b = 3 - b;
            if (a != b) { Lout = ii; return a > b ? 1 : -1; }
            if (a == 5) { Lout = ii; return -1; }
        }
    }
}

// suffixArraySearch1 with i1 = 0, i2 = nSA-1, L = 0
__device__ __forceinline__ u64 sjdbSearchOne(const SjdbIndex& ix, const u8* __restrict__ s) {
    u64 L1, L2, L3;
    if (sjdbCompare(ix, s, 0, 0, L1) < 0) return 0;
    if (sjdbCompare(ix, s, 0, ix.nSA - 1, L2) > 0) return ~0ULL - 1;   // behind the last row
    u64 i1 = 0, i2 = ix.nSA - 1, L = L1 < L2 ? L1 : L2;
#pragma unroll 1
    while (i1 + 1 < i2) {
        const u64 i3 = i1 / 2 + i2 / 2 + (i1 % 2 + i2 % 2) / 2;   // medianUint2
        if (sjdbCompare(ix, s, L, i3, L3) > 0) { i1 = i3; L1 = L3; }
        else { i2 = i3; L2 = L3; }
        L = L1 < L2 ? L1 : L2;
    }
    return i2;
}

__global__ void __launch_bounds__(256) sjdb_search_kernel(const SjdbIndex ix, const u8* __restrict__ Gsj, u64 nSeq, u64 sjdbLength,
                                                          const u8* __restrict__ skipSeq, u64* __restrict__ indArray) {
    const u64 nSuf = nSeq * sjdbLength;
#pragma unroll 1
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < nSuf; k += (u64)gridDim.x * blockDim.x) {
        const u64 q = k / sjdbLength;
        u64 row = ~0ULL;                                                     // no row: junction already in the index, or suffix starts with N / spacer
        if (!SB_LDG(skipSeq + q) && SB_LDG(Gsj + k) <= 3) row = sjdbSearchOne(ix, Gsj + k);
        indArray[2 * k] = row;
        indArray[2 * k + 1] = k;
    }
}

struct SjdbMerge {
    const u64* insRow;    // nInd: row of the NEW array every inserted suffix lands in (strictly increasing)
    const u64* insVal;    // nInd: its packed value
    u64 nInd, nSAnew;
    u64 nGenomeOld, nGenomeNew, sjGstart, sjdbLength, sjdbNold, nGsjNew;
    const u32* oldSJind;  // new index of old junction j
};

// an old row in the coordinates of the new index (sjdbBuildIndex.cpp:163-187)
__device__ __forceinline__ u64 sjdbRebase(const SjdbMerge& m, u64 ind1, u64 N2bit) {
    if (ind1 & N2bit) {
        u64 ind1s = m.nGenomeOld - (ind1 & ~N2bit);
        if (ind1s >= m.sjGstart) {   // inside an old insert: moves with its junction
            const u64 sj1 = (ind1s - m.sjGstart) / m.sjdbLength;
            if (sj1 < m.sjdbNold) ind1s += ((u64)SB_LDG(m.oldSJind + sj1) - sj1) * m.sjdbLength;
            ind1 = (m.nGenomeNew - ind1s) | N2bit;
        } else ind1 += m.nGsjNew;    // reverse-strand coordinates count from the (longer) end
    } else if (ind1 >= m.sjGstart) {
        const u64 sj1 = (ind1 - m.sjGstart) / m.sjdbLength;
        if (sj1 < m.sjdbNold) ind1 += ((u64)SB_LDG(m.oldSJind + sj1) - sj1) * m.sjdbLength;
    }
    return ind1;
}

// Tile = 32 groups of 64 rows per warp: every lane packs its group into the warp's shared-memory tile (same layout as the global
// array: group after group, `bits` words each), then the warp stores the tile with consecutive 8-byte words per lane (256 B per
// store instruction instead of 32 stores 8*bits bytes apart).  Dynamic shared memory: (blockDim.x/32) * 32 * bits * 8 bytes.
__global__ void __launch_bounds__(128) sjdb_merge_sa_kernel(const SjdbIndex ix, const SjdbMerge m, u64* __restrict__ SAnew) {
    extern __shared__ u8 smem[];
    const u32 bits = ix.saBits;
    const u32 lane = threadIdx.x & 31;
    u64* tile = (u64*)smem + (u64)(threadIdx.x >> 5) * 32 * bits;
    const u64 N2bit = 1ULL << ix.GstrandBit;
    const u64 nGroups = (m.nSAnew + 63) / 64;
    const u64 nWarps = ((u64)gridDim.x * blockDim.x) >> 5;
#pragma unroll 1
    for (u64 t = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t * 32 < nGroups; t += nWarps) {
        const u64 g = t * 32 + lane;
        if (g < nGroups) {
            const u64 r0 = g * 64;
            u64 lo = 0, hi = m.nInd;            // j = number of inserted rows in front of r0
            while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (SB_LDG(m.insRow + mid) < r0) lo = mid + 1; else hi = mid; }
            u64 j = lo;
            u64 nextIns = j < m.nInd ? SB_LDG(m.insRow + j) : ~0ULL;
            u64* out = tile + (u64)lane * bits;
            u64 acc = 0;
            u32 sh = 0;
#pragma unroll 1
            for (u32 e = 0; e < 64; e++) {
                const u64 r = r0 + e;
                u64 val = 0;
                if (r < m.nSAnew) {
                    if (r == nextIns) { val = SB_LDG(m.insVal + j); j++; nextIns = j < m.nInd ? SB_LDG(m.insRow + j) : ~0ULL; }
                    else val = sjdbRebase(m, packedGet(ix.SA, bits, r - j), N2bit);
                }
                acc |= val << sh;
                if (sh + bits >= 64) {
                    *out++ = acc;
                    acc = sh + bits > 64 ? val >> (64 - sh) : 0;
                }
                sh = (sh + bits) & 63;
            }
        }
        __syncwarp();
        const u64 rows = nGroups - t * 32 < 32 ? nGroups - t * 32 : 32;
        u64* dst = SAnew + t * 32 * bits;
        for (u64 k = lane; k < rows * bits; k += 32) dst[k] = tile[k];
        __syncwarp();
    }
}

// Host side of the merge launch: where every inserted suffix lands in the new array and what the row holds
// (sjdbBuildIndex.cpp:143-152, 191-199).  indSorted = nInd sorted (row, offset in Gsj) pairs; rows behind the last old one are appended.
inline void sjdbInsertedRows(const uint64_t* indSorted, u64 nInd, u64 nSAold, u64 nGsj, u64 sjGstart, u32 GstrandBit, u64* row, u64* val) {
    const u64 N2bit = 1ULL << GstrandBit;
    for (u64 j = 0; j < nInd; j++) {
        const u64 r = indSorted[2 * j] < nSAold ? indSorted[2 * j] : nSAold;
        row[j] = r + j;
        const u64 ind1 = indSorted[2 * j + 1];
        val[j] = ind1 < nGsj ? ind1 + sjGstart : ((ind1 - nGsj) | N2bit);
    }
}

}  // namespace starb
