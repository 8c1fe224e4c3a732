// sa_build_impl.cuh — suffix array of a genome on the GPU (reference source/Genome_genomeGenerate.cpp:178-330).
//
// The reference sorts 16-mer buckets of suffixes with qsort and an 8-bytes-at-a-time comparison (funCompareSuffixes, :29-89): hours
// for a mammalian genome.  Here: prefix doubling over ALL positions of the text T = G + reverse complement.
//   round 0   key = the first 20 codes of the suffix, 3 bits each, cut after the first code 5 ("terminated" key);
//             stable radix sort by key; rank = index of the first member of the run of equal keys, except that members of a
//             terminated run are all distinct and keep their (position) order: a 5 met at the same offset ends the reference's
//             comparison and the smaller text position goes first, which is exactly "every 5 is its own symbol, ordered by position".
//   round h   (h = 20, 40, 80, ...) key = (rank[p], rank[p+h]); sort; re-rank; until every rank is unique.
//   output    positions in rank order that hold a base (code < 4), packed at GstrandBit+1 bits (reverse-strand positions flagged).
// Every kernel is a grid-stride loop; the round loop below is written against a handful of macros (SA_LAUNCH, SA_ALLOC, ...) so that
// the same source runs on the device (sa_build.cu: CUDA + cub radix sort / scan / select) and, for tests, as emulated CTAs of host
// threads (oracle/engine_emul.cpp: the library primitives replaced by std:: equivalents).
// This path: 2*nGenome < 2^32 - 64 (32-bit ranks), ~37 bytes of HBM per text position; larger texts: sa_build_large.cuh.
#pragma once
#include "dev.cuh"

namespace starb {

#define SA_K0 20   /* codes in the round-0 key */

// T[i] = G[i], T[n-1-i] = complement, n = 2*nGenome; padT bytes of code 5 behind
__global__ void __launch_bounds__(256) sa_text_kernel(const u8* __restrict__ G, u64 nGenome, u8* __restrict__ T, u64 padT) {
    const u64 n = 2 * nGenome;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n + padT; i += (u64)gridDim.x * blockDim.x) {
        u8 c = 5;
        if (i < nGenome) c = SB_LDG(G + i);
        else if (i < n) { c = SB_LDG(G + (n - 1 - i)); if (c < 4) c = 3 - c; }
        T[i] = c;
    }
}

__global__ void __launch_bounds__(256) sa_key0_kernel(const u8* __restrict__ T, u64 n, u64* __restrict__ key, u32* __restrict__ pos) {
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        u64 k = 0;
        bool live = true;
#pragma unroll 1
        for (int j = 0; j < SA_K0; j++) {
            u64 c = 0;
            if (live) { c = SB_LDG(T + p + j); if (c == 5) live = false; }
            k = (k << 3) | c;
        }
        key[p] = k;
        pos[p] = (u32)p;
    }
}

__device__ __forceinline__ bool saHas5(u64 k) {   // some 3-bit digit of a round-0 key is 5 (binary 101)
    const u64 lo = 0x1249249249249249ULL;          // bit 0 of every digit
    return ((k >> 2) & ~(k >> 1) & k & lo) != 0;
}

// run heads of the sorted keys: hd[j] = j for the first member of a run (or any member of a terminated run in round 0), else 0
__global__ void __launch_bounds__(256) sa_heads_kernel(const u64* __restrict__ key, u64 n, int round0, u32* __restrict__ hd, unsigned long long* __restrict__ nHeads) {
    unsigned long long mine = 0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
        const u64 k = key[j];
        const bool head = j == 0 || k != key[j - 1] || (round0 && saHas5(k));
        hd[j] = head ? (u32)j : 0u;
        mine += head;
    }
    if (mine) atomicAdd(nHeads, mine);
}

// rank[pos[j]] = rankSorted[j]
__global__ void __launch_bounds__(256) sa_scatter_rank_kernel(const u32* __restrict__ pos, const u32* __restrict__ rankSorted, u64 n, u32* __restrict__ rank) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) rank[pos[j]] = rankSorted[j];
}

// doubling step: key = (rank[p], rank[p+h]); positions are taken in the current sorted order so that pos stays the permutation
__global__ void __launch_bounds__(256) sa_key_kernel(const u32* __restrict__ rank, const u32* __restrict__ pos, u64 n, u64 h, u64* __restrict__ key) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) {
        const u64 p = pos[j];
        const u64 q = p + h < n ? p + h : n;          // rank[n] = n: the end of the text sorts behind everything
        key[j] = ((u64)rank[p] << 32) | rank[q];
    }
}

__global__ void __launch_bounds__(256) sa_base_flag_kernel(const u8* __restrict__ T, const u32* __restrict__ pos, u64 n, u8* __restrict__ flag) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (u64)gridDim.x * blockDim.x) flag[j] = SB_LDG(T + pos[j]) < 4;
}

// one thread per 64 rows = `bits` whole words of the packed array; a warp's 32 groups are staged in shared memory and stored coalesced
// (dynamic shared memory: (blockDim.x/32) * 32 * bits * 8 bytes; launched with 128 threads)
__global__ void __launch_bounds__(128) sa_pack_kernel(const u32* __restrict__ sa, u64 nSA, u64 nGenome, u32 GstrandBit, u64* __restrict__ out) {
    extern __shared__ u8 smem[];
    const u32 bits = GstrandBit + 1;
    const u32 lane = threadIdx.x & 31;
    u64* tile = (u64*)smem + (u64)(threadIdx.x >> 5) * 32 * bits;
    const u64 N2bit = 1ULL << GstrandBit;
    const u64 nGroups = (nSA + 63) / 64;
    const u64 nWarps = ((u64)gridDim.x * blockDim.x) >> 5;
#pragma unroll 1
    for (u64 t = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t * 32 < nGroups; t += nWarps) {
        const u64 g = t * 32 + lane;
        if (g < nGroups) {
            u64* o = tile + (u64)lane * bits;
            u64 acc = 0;
            u32 sh = 0;
#pragma unroll 1
            for (u32 e = 0; e < 64; e++) {
                const u64 r = g * 64 + e;
                u64 val = 0;
                if (r < nSA) { const u64 p = sa[r]; val = p < nGenome ? p : ((p - nGenome) | N2bit); }
                acc |= val << sh;
                if (sh + bits >= 64) { *o++ = acc; acc = sh + bits > 64 ? val >> (64 - sh) : 0; }
                sh = (sh + bits) & 63;
            }
        }
        __syncwarp();
        const u64 rows = nGroups - t * 32 < 32 ? nGroups - t * 32 : 32;
        u64* dst = out + t * 32 * bits;
        for (u64 k = lane; k < rows * bits; k += 32) dst[k] = tile[k];
        __syncwarp();
    }
}

#ifdef SA_LAUNCH
// The build itself.  dG: device copy of G (nGenome bytes).  outWords: device buffer of ceil(nSA/64)*(GstrandBit+1) words.
// Returns 0, or a positive code: 1 = the number of bases differs from nSA, 2 = no convergence (cannot happen for a finite text).
inline int saBuildRun(const u8* dG, u64 nGenome, u32 GstrandBit, u64 nSA, u64* outWords, u64* roundsOut) {
    const u64 n = 2 * nGenome;
    const u64 padT = 64;
    u8* T = (u8*)SA_ALLOC(n + padT);
    u64* keyA = (u64*)SA_ALLOC(n * 8);
    u64* keyB = (u64*)SA_ALLOC(n * 8);
    u32* posA = (u32*)SA_ALLOC(n * 4);
    u32* posB = (u32*)SA_ALLOC(n * 4);
    u32* rank = (u32*)SA_ALLOC((n + 1) * 4);
    u32* hd = (u32*)SA_ALLOC(n * 4);
    unsigned long long* cnt = (unsigned long long*)SA_ALLOC(8);
    if (!T || !keyA || !keyB || !posA || !posB || !rank || !hd || !cnt) {
        SA_FREE(T); SA_FREE(keyA); SA_FREE(keyB); SA_FREE(posA); SA_FREE(posB); SA_FREE(rank); SA_FREE(hd); SA_FREE(cnt);
        return 3;
    }
    SA_LAUNCH(n + padT, sa_text_kernel, dG, nGenome, T, padT);
    SA_LAUNCH(n, sa_key0_kernel, T, n, keyA, posA);
    const u32 nU32 = (u32)n;
    SA_COPY_TO(rank + n, &nU32, 4);
    u64 rounds = 0;
    int rc = 2;
    for (u64 h = SA_K0;; h *= 2) {
        SA_SORT_PAIRS(keyA, keyB, posA, posB, n, rounds == 0 ? 3 * SA_K0 : 64);   // stable; sorted keys in keyB, positions in posB
        unsigned long long zero = 0, heads = 0;
        SA_COPY_TO(cnt, &zero, 8);
        SA_LAUNCH(n, sa_heads_kernel, keyB, n, rounds == 0 ? 1 : 0, hd, cnt);
        SA_MAX_SCAN(hd, n);                                                          // rank of every sorted row
        SA_LAUNCH(n, sa_scatter_rank_kernel, posB, hd, n, rank);
        SA_COPY_FROM(&heads, cnt, 8);
        rounds++;
        { u32* t = posA; posA = posB; posB = t; }                                   // posA = current order
        if (heads == n) { rc = 0; break; }
        if (h > 2 * n) break;
        SA_LAUNCH(n, sa_key_kernel, rank, posA, n, h, keyA);
    }
    if (rc == 0) {
        u8* flag = (u8*)keyB;                                                        // (keyB is free now)
        SA_LAUNCH(n, sa_base_flag_kernel, T, posA, n, flag);
        u64 nSel = 0;
        SA_SELECT(posA, flag, posB, n, &nSel);                                       // positions holding a base, in suffix order
        if (nSel != nSA) rc = 1;
        else SA_LAUNCH_PACK(nSA, GstrandBit + 1, sa_pack_kernel, posB, nSA, nGenome, GstrandBit, outWords);
    }
    SA_SYNC();
    if (roundsOut) *roundsOut = rounds;
    SA_FREE(T); SA_FREE(keyA); SA_FREE(keyB); SA_FREE(posA); SA_FREE(posB); SA_FREE(rank); SA_FREE(hd); SA_FREE(cnt);
    return rc;
}
#endif

}  // namespace starb
