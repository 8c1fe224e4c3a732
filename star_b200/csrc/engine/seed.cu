// seed.cu — read preparation + maximal-mappable-prefix (MMP) seed stage (sm_100a).
//
// What is computed (bit-exact with the reference, proven by tests/ against oracle/):
//   prep_reads_kernel    ReadAlign::oneRead prologue          reference source/ReadAlign_oneRead.cpp:35-78,
//                        convertNucleotidesToNumbers / complementSeqNumbers  SequenceFuns.cpp:4-14,131-146
//   seed stage           qualitySplit                          SequenceFuns.cpp:411-444
//                        ReadAlign::mapOneRead search schedule ReadAlign_mapOneRead.cpp:37-93
//                        maxMappableLength2strands             ReadAlign_maxMappableLength2strands.cpp:5-115
//                        maxMappableLength / findMultRange / compareSeqToGenome   SuffixArrayFuns.cpp:4-207
//                        storeAligns                           ReadAlign_storeAligns.cpp:10-160
// How (seed_keyed.cuh): seed_chains_kernel (chains of searches, one item each, keyed by SAindex L-mer) -> radix sort of the keys ->
// seed_keyed_search_kernel (groups of 8 lanes walk the chains: SAindex words, one coalesced load of the window's 32-bit SA keys, SA /
// genome only for the rows the keys leave undecided) -> seed_replay_kernel (records replayed through storeAligns in the reference's
// loop order).  The stage is HBM-sector bound; DESIGN.md states its algorithmic bytes and roofline.
// seed_search_warp_kernel (seed_warp.cuh: one read per warp, 32-ary search with genome comparisons, no keys) seeds the reads that are
// redone by the overflow tiers.
#include "dev.cuh"
#include "seed_warp.cuh"
#include "seed_keyed.cuh"

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

// One lane per read: the records of the read's searches, replayed through storeAligns in the reference's loop order
// (piece, direction, start, search; ReadAlign_mapOneRead.cpp:55-92): records carry (chainId, k) = that order.
__global__ void __launch_bounds__(128) seed_replay_kernel(const __grid_constant__ star_params_t P, ReadInfo* __restrict__ info, Piece* __restrict__ pieces, u32 maxP, u32 nReads,
                                                          const __grid_constant__ KeyedArgs ka) {
#pragma unroll 1
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nReads; i += gridDim.x * blockDim.x) {
        ReadInfo ri = info[i];
        StoreState st;
        st.PC = pieces + (u64)i * maxP;
        st.nP = 0; st.maxP = maxP; st.nA = 0; st.multNmin = 0; st.multNminL = 0; st.flags = ri.flags & 1u;
        const u32 n = ka.recCount[i];
        if (n > ka.maxRec) st.flags |= 1;                      // record slab full: the read is redone by the tier path
        u32 searches = 0, sai = 0;
        if (!st.flags) {
            const SeedRec* rec = ka.recs + (u64)i * ka.maxRec;
            if (n <= 256) {   // the usual case: sort (chainId, k, slot) words in a small local array, then replay in that order
                u32 ord[256];
#pragma unroll 1
                for (u32 r = 0; r < n; r++) {
                    const u32 key = ((u32)rec[r].chainId << 16) | ((u32)rec[r].k << 8);   // (chain ids use 12 bits; one record per (chain, k))
                    u32 j = r;
#pragma unroll 1
                    while (j > 0 && (ord[j - 1] >> 8) > (key >> 8)) { ord[j] = ord[j - 1]; j--; }
                    ord[j] = key | r;
                }
#pragma unroll 1
                for (u32 q = 0; q < n && !st.flags; q++) {
                    const SeedRec r = rec[ord[q] & 0xff];
                    searches++; sai += r.nSai;
                    storeAligns(st, P, (r.chainId >> 7) & 1u, r.Shift, r.Nrep, r.L, r.SAstart, r.iFrag);
                }
            } else {
                long long last = -1;
#pragma unroll 1
                for (u32 done = 0; done < n && !st.flags; done++) {
                    long long best = 1LL << 40;
                    u32 bi = 0;
#pragma unroll 1
                    for (u32 r = 0; r < n; r++) {
                        const long long key = ((long long)rec[r].chainId << 8) | rec[r].k;
                        if (key > last && key < best) { best = key; bi = r; }
                    }
                    last = best;
                    const SeedRec r = rec[bi];
                    searches++; sai += r.nSai;
                    storeAligns(st, P, (r.chainId >> 7) & 1u, r.Shift, r.Nrep, r.L, r.SAstart, r.iFrag);
                }
            }
        }
        ri.nP = (u16)st.nP;
        ri.nA = st.nA;
        ri.multNminL = st.multNminL;
        ri.flags = st.flags;
        ri.cSearches = searches; ri.cSaiWords = sai;
        ri.cSaEnum = 0; ri.cNodes = 0; ri.cLeaves = 0; ri.cSlow = 0;
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
void launch_build_sa_keys(int nSM, cudaStream_t stream, const DevIndex& ix, u32* keys) {
    build_sa_keys_kernel<<<nSM * 16, 256, 0, stream>>>(ix, keys);
}
void launch_seed_chains(int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info, u32 nReads, const KeyedArgs& ka) {
    seed_chains_kernel<<<nSM * 8, 128, 0, stream>>>(ix, P, reads, stride, info, nReads, ka);
}
template <u32 GN>
static void launchKeyedSearchG(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info, const u32* order,
                               const KeyedArgs& ka) {   // the occupancy target is part of the kernel (register budget): 8, 12 or 16 CTAs of 128 threads per SM
    if (ctasPerSM <= 8) seed_keyed_search_kernel<GN, 8><<<nSM * 8, 128, 0, stream>>>(ix, P, reads, stride, info, order, ka);
    else if (ctasPerSM <= 12) seed_keyed_search_kernel<GN, 12><<<nSM * 12, 128, 0, stream>>>(ix, P, reads, stride, info, order, ka);
    else seed_keyed_search_kernel<GN, 16><<<nSM * 16, 128, 0, stream>>>(ix, P, reads, stride, info, order, ka);
}
void launch_seed_keyed_search(int groupLanes, int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info,
                              const u32* order, const KeyedArgs& ka) {
    if (groupLanes == 4) launchKeyedSearchG<4>(ctasPerSM, nSM, stream, ix, P, reads, stride, info, order, ka);
    else if (groupLanes == 16) launchKeyedSearchG<16>(ctasPerSM, nSM, stream, ix, P, reads, stride, info, order, ka);
    else launchKeyedSearchG<8>(ctasPerSM, nSM, stream, ix, P, reads, stride, info, order, ka);
}
void launch_seed_replay(int nSM, cudaStream_t stream, const star_params_t& P, ReadInfo* info, Piece* pieces, u32 maxP, u32 nReads, const KeyedArgs& ka) {
    seed_replay_kernel<<<nSM * 8, 128, 0, stream>>>(P, info, pieces, maxP, nReads, ka);
}
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
