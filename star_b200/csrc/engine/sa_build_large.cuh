// sa_build_large.cuh — suffix-array build for texts beyond 32-bit ranks / beyond one sort of the whole text (GRCh38: n = 2 x 3.1 G).
//
// Same ordering as sa_build_impl.cuh (prefix doubling, every code 5 its own symbol ordered by position), arranged so that no sort is
// larger than `cap` elements and only 64-bit rank / order arrays of the whole text stay resident:
//   round 0   the text is split by the first 4 codes (4096 bins = the top 12 bits of the 60-bit key) into runs of bins of <= cap
//             positions; every run is selected (in position order), keyed, sorted and ranked on its own — bins are in key order, so
//             the runs concatenate to the global order.
//   round h   only members of groups that are still tied are touched ("active"): batches of consecutive groups (<= cap order slots,
//             cut at a group start) are selected, keyed with (group start - batch start, rank[p+h]) in ONE 64-bit word (31 + 33 bits),
//             sorted, and written back into the group's own slots with refined ranks.  Ranks are refined in place: a later batch may
//             read ranks already refined in this round, which only sharpens its keys (rank[a] < rank[b] always implies suffix a <
//             suffix b).  Stops when a round finds nothing active.
//   heads     one byte per order slot says "a group starts here"; it is written by the same kernels that write the order, so a round finds
//             its active slots (members of groups with more than one member) by streaming that array — no gather through rank[order[j]].
//   output    the first code is the most significant digit of every key, so the positions holding a base (codes 0..3) are exactly the first
//             nSA order slots: they are packed straight from `order`, nothing is compacted.
// HBM at n positions: text n + heads n + order 8n + rank 8n + 4 x 8 x cap work buffers + sort scratch (~2 x 8 x cap); the genome copy is
// released once the text exists and the packed output is allocated after the ranks are released (GRCh38 with cap = 7e8: ~146 GB at the peak).
// Limits: n < 2^33, cap <= 2^31, no single 4-mer bin and no single tied group larger than cap.
// Written against the SA_* macros of sa_build_impl.cuh plus SA_SELECT_IF / SA_MAX_SCAN64 / SA_SORT_PAIRS64.
#pragma once
#include <vector>

#include "sa_build_impl.cuh"

namespace starb {

struct SaBinInRange {   // position p starts with a 4-code prefix whose bin is in [b0, b1)
    const u8* T; u32 b0, b1;
    __host__ __device__ bool operator()(u64 p) const {
        u32 b = 0; bool live = true;
        for (int j = 0; j < 4; j++) { u32 c = 0; if (live) { c = T[p + j]; if (c == 5) live = false; } b = (b << 3) | c; }
        return b >= b0 && b < b1;
    }
};
struct SaActive {       // order slot j belongs to a group of more than one member (head[j]: a group starts at slot j; head[n] = 1)
    const u8* head;
    __host__ __device__ bool operator()(u64 j) const { return !(head[j] && head[j + 1]); }
};

__global__ void __launch_bounds__(256) sal_hist_kernel(const u8* __restrict__ T, u64 n, unsigned long long* __restrict__ hist) {
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x) {
        u32 b = 0; bool live = true;
        for (int j = 0; j < 4; j++) { u32 c = 0; if (live) { c = SB_LDG(T + p + j); if (c == 5) live = false; } b = (b << 3) | c; }
        atomicAdd(hist + b, 1ULL);
    }
}
__global__ void __launch_bounds__(256) sal_key0_kernel(const u8* __restrict__ T, const u64* __restrict__ pos, u64 m, u64* __restrict__ key) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 p = pos[i];
        u64 k = 0; bool live = true;
#pragma unroll 1
        for (int j = 0; j < SA_K0; j++) { u64 c = 0; if (live) { c = SB_LDG(T + p + j); if (c == 5) live = false; } k = (k << 3) | c; }
        key[i] = k;
    }
}
// run heads of sorted round-0 keys, as (local index + 1) for heads and 0 otherwise (index 0 is always a head, so max-scan works with +1)
__global__ void __launch_bounds__(256) sal_heads0_kernel(const u64* __restrict__ key, u64 m, u64* __restrict__ hd) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = key[i];
        hd[i] = (i == 0 || k != key[i - 1] || saHas5(k)) ? i + 1 : 0;
    }
}
// (hd after the max-scan: index + 1 of the latest head at or before i; i is a head itself iff hd[i] == i + 1)
__global__ void __launch_bounds__(256) sal_write0_kernel(const u64* __restrict__ pos, const u64* __restrict__ hd, u64 m, u64 base, u64* __restrict__ order, u64* __restrict__ rank,
                                                         u8* __restrict__ head) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        order[base + i] = pos[i];
        rank[pos[i]] = base + hd[i] - 1;
        head[base + i] = hd[i] == i + 1;
    }
}
__global__ void __launch_bounds__(256) sal_key_kernel(const u64* __restrict__ rank, const u64* __restrict__ order, const u64* __restrict__ slot, u64 m, u64 a, u64 h, u64 n,
                                                      u64* __restrict__ key, u64* __restrict__ val) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 p = order[slot[i]];
        const u64 q = p + h < n ? p + h : n;
        key[i] = ((rank[p] - a) << 33) | rank[q];
        val[i] = p;
    }
}
// gh: (index + 1) of the first list member of the element's group; nh: (index + 1) of the latest member that starts a new sub-group
__global__ void __launch_bounds__(256) sal_heads_kernel(const u64* __restrict__ key, u64 m, u64* __restrict__ gh, u64* __restrict__ nh) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = key[i];
        const bool g = i == 0 || (k >> 33) != (key[i - 1] >> 33);
        gh[i] = g ? i + 1 : 0;
        nh[i] = (g || k != key[i - 1]) ? i + 1 : 0;
    }
}
__global__ void __launch_bounds__(256) sal_write_kernel(const u64* __restrict__ key, const u64* __restrict__ val, const u64* __restrict__ gh, const u64* __restrict__ nh, u64 m, u64 a,
                                                        u64* __restrict__ order, u64* __restrict__ rank, u8* __restrict__ head) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 r1 = a + (key[i] >> 33);
        const u64 slot = r1 + (i + 1 - gh[i]);
        order[slot] = val[i];
        rank[val[i]] = r1 + (nh[i] - gh[i]);
        head[slot] = nh[i] == i + 1;     // this member opens a (sub-)group: the old group's first slot stays a head, new ones appear inside it
    }
}
// as sa_pack_kernel, 64-bit positions
__global__ void __launch_bounds__(128) sal_pack_kernel(const u64* __restrict__ sa, u64 nSA, u64 nGenome, u32 GstrandBit, u64* __restrict__ out) {
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
// dG: device copy of G, RELEASED here as soon as the text exists.  *outWordsPtr: allocated here (nOutWords zeroed words) once the ranks
// are released; the caller frees it.  Returns 0, or: 1 = number of bases differs from nSA, 2 = no convergence, 3 = out of memory,
// 4 = a 4-mer bin or a tied group exceeds cap.
inline int saBuildRunLarge(u8* dG, u64 nGenome, u32 GstrandBit, u64 nSA, u64** outWordsPtr, u64 nOutWords, u64 cap, u64* roundsOut) {
    const u64 n = 2 * nGenome, padT = 64;
    *outWordsPtr = nullptr;
    if (n >= (1ULL << 33) || cap > (1ULL << 31) || cap < 64) { SA_FREE(dG); return 4; }
    u8* T = (u8*)SA_ALLOC(n + padT);
    if (T) SA_LAUNCH(n + padT, sa_text_kernel, dG, nGenome, T, padT);
    SA_SYNC();
    SA_FREE(dG);
    u8* head = (u8*)SA_ALLOC(n + 1);
    u64* rank = (u64*)SA_ALLOC((n + 1) * 8);
    u64* order = (u64*)SA_ALLOC(n * 8);
    u64 *keyA = (u64*)SA_ALLOC(cap * 8), *keyB = (u64*)SA_ALLOC(cap * 8), *valA = (u64*)SA_ALLOC(cap * 8), *valB = (u64*)SA_ALLOC(cap * 8);
    unsigned long long* hist = (unsigned long long*)SA_ALLOC(4096 * 8);
    int rc = 0;
    u64 rounds = 0;
    if (!T || !head || !rank || !order || !keyA || !keyB || !valA || !valB || !hist) rc = 3;
    if (!rc) {
        SA_COPY_TO(rank + n, &n, 8);
        const u8 one = 1;
        SA_COPY_TO(head + n, &one, 1);
        std::vector<unsigned long long> h4(4096, 0);
        SA_COPY_TO(hist, h4.data(), 4096 * 8);
        SA_LAUNCH(n, sal_hist_kernel, T, n, hist);
        SA_COPY_FROM(h4.data(), hist, 4096 * 8);
        {   // positions that start with a base = bins whose first digit is 0..3 = the first nSA slots of the final order
            u64 nBase = 0;
            for (u32 b = 0; b < 4096; b++) if ((b >> 9) < 4) nBase += h4[b];
            if (nBase != nSA) rc = 1;
        }
        // ---- round 0: runs of bins
        u64 base = 0;
        for (u32 b0 = 0; b0 < 4096 && !rc;) {
            u64 cnt = h4[b0];
            u32 b1 = b0 + 1;
            if (cnt > cap) { rc = 4; break; }
            while (b1 < 4096 && cnt + h4[b1] <= cap) cnt += h4[b1++];
            if (cnt) {
                u64 m = 0;
                SA_SELECT_IF((SaBinInRange{T, b0, b1}), 0, n, valA, &m);
                if (m != cnt) { rc = 2; break; }
                SA_LAUNCH(m, sal_key0_kernel, T, valA, m, keyA);
                SA_SORT_PAIRS64(keyA, keyB, valA, valB, m, 3 * SA_K0);
                SA_LAUNCH(m, sal_heads0_kernel, keyB, m, keyA);
                SA_MAX_SCAN64(keyA, m);
                SA_LAUNCH(m, sal_write0_kernel, valB, keyA, m, base, order, rank, head);
                base += m;
            }
            b0 = b1;
        }
        if (!rc && base != n) rc = 2;
        rounds = 1;
        // ---- doubling rounds over the active groups
        for (u64 h = SA_K0; !rc; h *= 2) {
            u64 active = 0;
            for (u64 a = 0; a < n && !rc;) {
                u64 b = a + cap < n ? a + cap : n;
                if (b < n) {   // cut at the start of the group that slot b belongs to
                    u64 pb = 0, gs = 0;
                    SA_COPY_FROM(&pb, order + b, 8);
                    SA_COPY_FROM(&gs, rank + pb, 8);
                    if (gs <= a) { rc = 4; break; }
                    b = gs;
                }
                u64 m = 0;
                SA_SELECT_IF((SaActive{head}), a, b, valA, &m);   // active slots of the batch
                if (m) {
                    SA_LAUNCH(m, sal_key_kernel, rank, order, valA, m, a, h, n, keyA, valB);
                    SA_SORT_PAIRS64(keyA, keyB, valB, valA, m, 64);           // sorted keys in keyB, positions in valA
                    SA_LAUNCH(m, sal_heads_kernel, keyB, m, keyA, valB);
                    SA_MAX_SCAN64(keyA, m);
                    SA_MAX_SCAN64(valB, m);
                    SA_LAUNCH(m, sal_write_kernel, keyB, valA, keyA, valB, m, a, order, rank, head);
                    active += m;
                }
                a = b;
            }
            if (rc) break;
            if (active == 0) break;
            rounds++;
            if (h > 2 * n) { rc = 2; break; }
        }
    }
    SA_SYNC();
    SA_FREE(keyA); SA_FREE(keyB); SA_FREE(valA); SA_FREE(valB); SA_FREE(hist); SA_FREE(rank); SA_FREE(head); SA_FREE(T);   // (the order is final)
    if (!rc) {
        u64* outWords = (u64*)SA_ALLOC(nOutWords * 8);
        if (!outWords) rc = 3;
        else {
            SA_ZERO(outWords, nOutWords * 8);
            SA_LAUNCH_PACK(nSA, GstrandBit + 1, sal_pack_kernel, order, nSA, nGenome, GstrandBit, outWords);
            *outWordsPtr = outWords;
        }
    }
    SA_SYNC();
    if (roundsOut) *roundsOut = rounds;
    SA_FREE(order);
    return rc;
}
#endif

}  // namespace starb
