// seed_keyed.cuh — the default maximal-mappable-prefix (MMP) seed stage: keyed SA windows, chains binned by SAindex L-mer.
//
// Same stored pieces as the reference (ReadAlign_mapOneRead.cpp:37-93, ReadAlign_maxMappableLength2strands.cpp:5-115,
// SuffixArrayFuns.cpp:10-207, ReadAlign_storeAligns.cpp:10-160), organised for the B200 memory system instead of a CPU cache:
//
//  * SA keys.  At context creation one pass over the suffix array writes a 32-bit key per SA row: the 14 bases that FOLLOW the
//    SAindex prefix of that suffix (2 bits each, in comparison orientation: reverse-strand rows complemented) and the number of
//    valid bases before an N / padding / junction spacer.  5.9 G rows x 4 B = 23.6 GB for GRCh38 — HBM the CPU layout cannot afford
//    and a B200 has.  All rows of a SAindex interval share the prefix, so the answer of maxMappableLength over such a window — the
//    maximal match length and the block of rows attaining it — is decided by the keys for every row that mismatches within the next
//    14 bases: ONE coalesced load of the window's keys replaces the ~7.5 dependent (SA word -> genome bytes) probes of the binary
//    search.  Only the rows that match all 28 bases (normally one: the locus the read came from) touch SA and genome.
//  * Chains.  mapOneRead's searches form independent chains (piece x direction x start); a chain's searches are sequential
//    (the next one starts where the previous match ended), chains are not.  A chunk's chains (~10 per pair) become one item list,
//    sorted by the SAindex L-mer of their first search, and are walked by groups of 8 lanes: neighbouring groups read neighbouring
//    SAindex words and key windows.  Every search leaves a 24-byte record in its read's slab.
//  * Replay.  storeAligns is order-dependent (first of two equal pieces wins, multNmin tracking, caps), so one lane per read replays
//    the records in the reference's loop order (piece, direction, start, search) through the unchanged storeAligns.
//
// A window that is not a plain SAindex interval (prefix contains N, last L-mer of a length, piece shorter than the prefix) and the rows
// left undecided by the keys go through a k-ary search over the rows with genome comparisons (groupMaxMappable: the order of the probes
// does not change the answer — checked against the reference's binary search on millions of searches by the oracle).
#pragma once
#include "dev.cuh"
#include "seed_types.cuh"
#include "seed_warp.cuh"

namespace starb {

#define SK_KEY_BASES 14u

// ---- groups of Grp::G lanes inside a warp: every collective names the group's own lanes, so groups advance independently
template <u32 GN>
struct GrpT {
    static constexpr u32 G = GN;            // lanes per group
    static constexpr u32 TILE = 4u * GN;    // SA rows per key tile: one 16-byte load per lane
    u32 lane, shift, mask;
    SB_DEV GrpT() {
        const u32 l = threadIdx.x & 31;
        lane = l & (G - 1); shift = l & ~(G - 1); mask = ((1u << G) - 1u) << shift;
    }
    SB_DEV u32 ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & ((1u << G) - 1u); }
    SB_DEV u32 shfl(u32 v, int src) const { return __shfl_sync(mask, v, src, G); }
    SB_DEV u64 shfl64(u64 v, int src) const { return __shfl_sync(mask, v, src, G); }
    SB_DEV u32 maxU(u32 v) const { for (u32 o = G / 2; o; o >>= 1) { const u32 x = __shfl_xor_sync(mask, v, o, G); v = x > v ? x : v; } return v; }
    SB_DEV u32 minU(u32 v) const { for (u32 o = G / 2; o; o >>= 1) { const u32 x = __shfl_xor_sync(mask, v, o, G); v = x < v ? x : v; } return v; }
    SB_DEV void sync() const { __syncwarp(mask); }
};

// key of one suffix: 14 bases from offset Lk in comparison orientation, first base in bits 31:30; bits 3:0 = 14 - (valid bases before the
// first code > 3); the slots behind an invalid base are filled with 3 so that keys are ordered like the suffixes (N sorts last)
SB_DEV u32 saKeyOfRow(const DevIndex& ix, u64 iSA) {
    u64 SAstr = packedGet(ix.SA, ix.saBits, iSA);
    const bool dirG = (SAstr >> ix.GstrandBit) == 0;
    SAstr &= ix.GstrandMask;
    const u32 Lk = ix.gSAindexNbases;
    u64 w0, w1;
    if (dirG) {
        const u8* g = ix.G + SAstr + Lk;
        w0 = load8global(g); w1 = load8global(g + 8);
    } else {
        const u8* g = ix.G + (ix.nGenome - 1 - SAstr - Lk);
        w0 = bswap64(load8global(g - 7)); w1 = bswap64(load8global(g - 15));
    }
    u32 key = 0, nv = SK_KEY_BASES;
#pragma unroll
    for (u32 k = 0; k < SK_KEY_BASES; k++) {
        u32 c = (u32)((k < 8 ? w0 >> (8 * k) : w1 >> (8 * (k - 8))) & 0xff);
        if (nv == SK_KEY_BASES && c > 3) nv = k;
        if (nv != SK_KEY_BASES) c = 3;
        else if (!dirG) c = 3 - c;
        key |= c << (30 - 2 * k);
    }
    return key | (SK_KEY_BASES - nv);
}

__global__ void __launch_bounds__(256) build_sa_keys_kernel(const __grid_constant__ DevIndex ix, u32* __restrict__ keys) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ix.nSA; i += (u64)gridDim.x * blockDim.x) keys[i] = saKeyOfRow(ix, i);
}

// 8 codes 0..3, one per byte, byte 7 first -> 16 bits, 2 per code (the first code in bits 15:14)
SB_DEV u64 pack8(u64 w) {
    u64 x = w & 0x0303030303030303ULL;
    x = (x | (x >> 6)) & 0x000F000F000F000FULL;
    x = (x | (x >> 12)) & 0x000000FF000000FFULL;
    x = (x | (x >> 24)) & 0xFFFFULL;
    return x;
}

// base ii of the piece in comparison orientation (the piece holds codes 0..3 only)
SB_DEV u32 pieceBase(const u8* R, u64 S, bool dirR, u32 ii) { return dirR ? (u32)R[S + ii] : 3u - (u32)R[S - ii]; }

// match length (from offset L on) of the piece with ONE row, the group's lanes comparing 8 bases each per step.  Returns the length, N = all.
template <class Grp>
SB_DEV u32 grpLcpRow(const Grp& g, const DevIndex& ix, const u8* R, u64 S, u32 N, u32 L, u64 iSA, bool dirR, u32& bases) {
    u64 SAstr = packedGet(ix.SA, ix.saBits, iSA);
    const bool dirG = (SAstr >> ix.GstrandBit) == 0;
    SAstr &= ix.GstrandMask;
    const bool compl_ = dirR != dirG;
    const u8* g0 = ix.G + (dirG ? (long long)(SAstr + L) : (long long)(ix.nGenome - 1 - SAstr - L));
    const u8* r0 = R + (dirR ? (long long)(S + L) : (long long)S - (long long)L);
    const u32 n = N - L;
    u32 res = n;
#pragma unroll 1
    for (u32 base = 0; base < n; base += 8 * Grp::G) {
        const u32 ii = base + 8 * g.lane;
        u32 mine = 0xffffffffu;                          // first mismatch seen by this lane (offset from L), none: big
        if (ii < n) {
            u64 rs = dirR ? load8generic(r0 + ii) : bswap64(load8generic(r0 - (long long)ii - 7));
            const u64 gs = dirG ? load8global(g0 + ii) : bswap64(load8global(g0 - (long long)ii - 7));
            if (compl_) rs ^= 0x0303030303030303ULL;
            const u64 x = rs ^ gs;
            if (x) { const u32 k = (u32)SB_CTZ64(x) >> 3; if (ii + k < n) mine = ii + k; }
        }
        const u32 first = g.minU(mine);
        if (first != 0xffffffffu) { res = first; break; }
    }
    if (g.lane == 0) bases += res < n ? res + 1 : n;
    return L + res;
}

// The block [b1,b2] of SA rows of [lo,hi] whose match length with the piece is maximal, and that length; rows of [lo,hi] share the first
// Lc bases with the piece, L is the length the caller's interval guarantees (as in warpMaxMappableLength, with Grp::G probes per step).
struct BlockRes { u64 b1, b2; u32 L, probes, bases; };   // returned BY VALUE by the out-of-line searches: taking the address of a hot
                                                          // variable (probes, bases, maxL) would move it to local memory for the whole kernel
template <class Grp>
__device__ __noinline__ BlockRes groupMaxMappable(const Grp& g, const DevIndex& ix, const u8* R, u64 S, u32 N, u64 lo, u64 hi, bool dirR, u32 Lin) {
    u32 probes = 0, bases = 0;
    const u32 lane = g.lane;
    u64 i1 = lo, i2 = hi, i3 = lo;
    u32 L3 = 0, Lc = Lin;
    bool have = false;
#pragma unroll 1
    while (i2 - i1 + 1 > Grp::G) {
        const u64 row = i1 + ((i2 - i1) * lane) / (Grp::G - 1);   // (rows < 2^34: the product fits 64 bits)
        bool c;
        const u32 Lj = lcpRow<Grp>(ix, R, S, N, Lc, row, dirR, c, bases);
        probes += Grp::G;
        const u32 fullMask = g.ballot(Lj == N);
        const u32 gtMask = g.ballot(Lj != N && c);
        if (fullMask) { const int src = SB_FFS(fullMask) - 1; i3 = g.shfl64(row, src); L3 = N; have = true; break; }
        if (!gtMask) { i3 = g.shfl64(row, 0); L3 = g.shfl(Lj, 0); have = true; break; }
        const int jLast = 31 - SB_CLZ(gtMask);
        if (jLast == (int)Grp::G - 1) { i3 = g.shfl64(row, Grp::G - 1); L3 = g.shfl(Lj, Grp::G - 1); have = true; break; }
        const u64 n1 = g.shfl64(row, jLast), n2 = g.shfl64(row, jLast + 1);
        const u32 l1 = g.shfl(Lj, jLast), l2 = g.shfl(Lj, jLast + 1);
        i1 = n1; i2 = n2;
        Lc = l1 < l2 ? l1 : l2;
    }
    u64 b1, b2;
    bool seenLeft = false, seenRight = false;
    if (!have) {
        const u32 rows = (u32)(i2 - i1 + 1);
        const bool valid = lane < rows;
        bool c = false;
        const u32 Lj = valid ? lcpRow<Grp>(ix, R, S, N, Lc, i1 + lane, dirR, c, bases) : 0;
        probes += rows;
        L3 = g.maxU(valid ? Lj + 1 : 0) - 1;
        const u32 eq = g.ballot(valid && Lj == L3);
        const int jm = SB_FFS(eq) - 1;
        const u32 run = eq >> jm;
        const int len = SB_CTZ(~run);                    // (eq has at most Grp::G < 32 bits: ~run is never 0)
        const int j1 = jm, j2 = jm + len - 1;
        b1 = i1 + (u64)j1; b2 = i1 + (u64)j2;
        i3 = b1;
        seenLeft = j1 > 0; seenRight = (u32)j2 < rows - 1;
    } else {
        b1 = b2 = i3;
    }
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const bool left = side == 0;
        if (left ? (seenLeft || b1 <= lo) : (seenRight || b2 >= hi)) continue;
        const u64 inRow = left ? b1 : b2, outRow = left ? lo : hi;
        bool c;
        const u32 LoutRow = lcpRow<Grp>(ix, R, S, L3, Lin, outRow, dirR, c, bases);
        probes += 1;
        u64 res;
        if (LoutRow >= L3) res = outRow;
        else {
            u64 a = outRow, b = inRow;
            u32 La = LoutRow;
#pragma unroll 1
            while (left ? a + 1 < b : b + 1 < a) {
                const u64 span = (left ? b - a : a - b) - 1;
                const u32 np = span < Grp::G ? (u32)span : Grp::G;
                const bool valid = lane < np;
                const u64 step = np == span ? 1 + lane : ((span + 1) * (lane + 1)) / (np + 1);
                const u64 row = left ? a + step : a - step;
                const u32 Lj = valid ? lcpRow<Grp>(ix, R, S, L3, La, row, dirR, c, bases) : 0;
                probes += np;
                const u32 inMask = g.ballot(valid && Lj >= L3);
                if (!inMask) { a = g.shfl64(row, (int)np - 1); La = g.shfl(Lj, (int)np - 1); }
                else {
                    const int jIn = SB_FFS(inMask) - 1;
                    b = g.shfl64(row, jIn);
                    if (jIn > 0) { a = g.shfl64(row, jIn - 1); La = g.shfl(Lj, jIn - 1); }
                }
            }
            res = b;
        }
        if (left) b1 = res; else b2 = res;
    }
    BlockRes r;
    r.b1 = b1; r.b2 = b2; r.L = L3; r.probes = probes; r.bases = bases;
    return r;
}

// match length of a row's key with the piece's key over m bases (both left-aligned in 28 bits)
SB_DEV u32 keyLcp(u32 key, u32 rk, u32 m) {
    const u32 x = ((key >> 4) ^ rk) << 4;
    const u32 lead = x ? (u32)SB_CLZ(x) >> 1 : SK_KEY_BASES;
    const u32 nv = SK_KEY_BASES - (key & 15u);
    u32 l = lead < nv ? lead : nv;
    return l < m ? l : m;
}
// order of a row's key relative to the piece's key over the first p bases: -1 row < piece, 0 equal, +1 row > piece (an N is larger than any base)
SB_DEV int keyCmp(u32 key, u32 rk, u32 p) {
    const u32 sh = 4 + 2 * (SK_KEY_BASES - p);
    const u32 a = key >> sh, b = (rk << 4) >> sh;
    if (a != b) return a < b ? -1 : 1;
    return (SK_KEY_BASES - (key & 15u)) < p ? 1 : 0;
}

// first row of [i1, i2] whose key over p bases is >= (upper: >) the piece's; i2 + 1 if none
__device__ __noinline__ u64 keyBound(const u32* __restrict__ keys, u64 i1, u64 i2, u32 rk, u32 p, bool upper, u32& probes) {
    u64 a = i1, b = i2 + 1;
    while (a < b) {
        const u64 mid = a + (b - a) / 2;
        const int c = keyCmp(SB_LDG(keys + mid), rk, p);
        probes++;
        if (upper ? c <= 0 : c < 0) a = mid + 1; else b = mid;
    }
    return a;
}
// keyedWindow for a window too large to scan: the rows equal to the piece over m bases, else the block around the insertion point that
// shares the most bases with it (the maximal match is attained next to the insertion point; its block = the rows equal over that many bases).
// Every lane of the group runs the same bisection (broadcast loads).
__device__ __noinline__ BlockRes keyedWindowBisect(const u32* __restrict__ keys, u64 i1, u64 i2, u32 rk, u32 m) {
    u32 probes = 0;
    BlockRes r;
    r.bases = 0;
    const u64 lb = keyBound(keys, i1, i2, rk, m, false, probes), ub = keyBound(keys, i1, i2, rk, m, true, probes);
    if (ub > lb) { r.b1 = lb; r.b2 = ub - 1; r.L = m; r.probes = probes; return r; }
    u32 la = 0, lbv = 0;                                   // neighbours of the insertion point
    if (lb > i1) la = keyLcp(SB_LDG(keys + lb - 1), rk, m);
    if (lb <= i2) lbv = keyLcp(SB_LDG(keys + lb), rk, m);
    const u32 best = la > lbv ? la : lbv;
    r.b1 = best ? keyBound(keys, i1, i2, rk, best, false, probes) : i1;
    r.b2 = best ? keyBound(keys, i1, i2, rk, best, true, probes) - 1 : i2;
    r.L = best; r.probes = probes;
    return r;
}

// Keyed window: rows [i1,i2] share the SAindex prefix (Lk bases) with the piece.  Finds the maximal match length over the next m = min(14, N-Lk)
// bases and the block of rows attaining it from the keys alone.  Returns that length (0..m).
template <class Grp>
SB_DEV u32 keyedWindow(const Grp& g, const u32* __restrict__ keys, u32 scanMax, u64 i1, u64 i2, u32 rk, u32 m, u64& b1, u64& b2, u32& probes) {
    const u64 rows = i2 - i1 + 1;
    if (rows <= scanMax) {
        u32 best = 0;
        u64 first = i1, last = i1;
        bool any = false;
#pragma unroll 1
        for (u64 t = i1 & ~3ULL; t <= i2; t += Grp::TILE) {
            const u64 r0 = t + 4 * g.lane;
            uint4 kv = make_uint4(0, 0, 0, 0);
            if (r0 <= i2) kv = SB_LDG((const uint4*)(keys + r0));
            const u32 kk[4] = {kv.x, kv.y, kv.z, kv.w};
            u32 lmax = 0, lo = 0xffffffffu, hi = 0;
#pragma unroll
            for (u32 s = 0; s < 4; s++) {
                const u64 r = r0 + s;
                if (r >= i1 && r <= i2) {
                    const u32 l = keyLcp(kk[s], rk, m) + 1;            // +1: 0 = no row
                    if (l > lmax) { lmax = l; lo = 4 * g.lane + s; hi = lo; }
                    else if (l == lmax) hi = 4 * g.lane + s;
                }
            }
            const u32 tmax = g.maxU(lmax);
            if (tmax == 0) continue;
            const u32 tf = g.minU(lmax == tmax ? lo : 0xffffffffu), tl = g.maxU(lmax == tmax ? hi + 1 : 0) - 1;
            if (!any || tmax - 1 > best) { best = tmax - 1; first = t + tf; last = t + tl; any = true; }
            else if (tmax - 1 == best) last = t + tl;
        }
        probes += (u32)rows;
        b1 = first; b2 = last;
        return best;
    }
    const BlockRes r = keyedWindowBisect(keys, i1, i2, rk, m);   // large window (repeats, low-complexity prefixes): cold, out of line
    b1 = r.b1; b2 = r.b2; probes += r.probes;
    return r.L;
}

// ReadAlign_maxMappableLength2strands.cpp:5-115 (gSAsparseD == 1) for one piece by one group.  Returns the record fields.
// plain: the searched bases are codes 0..3 (always, except the reverse --seedSearchLmax search, which may reach in front of its piece:
// such a search keeps the reference's arithmetic on the codes > 3 and never uses the keys).
template <class Grp>
SB_DEV void groupSearch(const Grp& g, const DevIndex& ix, const u32* __restrict__ keys, u32 scanMax, const u8* R, u64 pieceStart, u32 pieceLength, bool dirR, bool plain,
                        u32& maxLout, u64& Nrep, u64& SAstart, u32& nSai, u32& probes, u32& bases) {
    u64 indStartEnd[2] = {0, 0};
    const u32 Lmax = ix.gSAindexNbases < pieceLength ? ix.gSAindexNbases : pieceLength;
    u64 ind1 = 0;
    u32 rk28 = 0;   // the 14 bases behind the SAindex prefix (2 bits each, left-aligned in 28 bits): what the SA keys are compared with
    if (plain) {
        // 32 bases of the piece in comparison orientation from four 8-byte gathers, packed 2 bits per base (base 0 in the top bits);
        // the bases behind the piece are never used
        u64 w0, w1, w2, w3;
        const u8* p = R + pieceStart;
        if (dirR) { w0 = bswap64(load8generic(p)); w1 = bswap64(load8generic(p + 8)); w2 = bswap64(load8generic(p + 16)); w3 = bswap64(load8generic(p + 24)); }
        else {   // (base k = 3 - R[S - k]: byte 7 of the word at S-7 is base 0 already)
            const u64 c3 = 0x0303030303030303ULL;
            w0 = load8generic(p - 7) ^ c3; w1 = load8generic(p - 15) ^ c3; w2 = load8generic(p - 23) ^ c3; w3 = load8generic(p - 31) ^ c3;
        }
        const u64 pb = (pack8(w0) << 48) | (pack8(w1) << 32) | (pack8(w2) << 16) | pack8(w3);
        ind1 = Lmax ? pb >> (64 - 2 * Lmax) : 0;
        rk28 = (u32)((pb << (2 * ix.gSAindexNbases)) >> 36);
    } else {
#pragma unroll 1
        for (u32 ii = 0; ii < Lmax; ii++) ind1 = (ind1 << 2) + (dirR ? (u64)R[pieceStart + ii] : 3 - (u64)R[pieceStart - ii]);   // (64-bit, as the reference: ReadAlign_maxMappableLength2strands.cpp:33-36)
    }
    u32 Lind = Lmax;
    u64 iSA1 = 0, iSA2 = 0;
    nSai = 0;
#pragma unroll 1
    while (Lind > 0) {
        iSA1 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1);
        nSai++;
        if ((iSA1 & ix.SAiMarkAbsentMaskC) == 0) break;
        --Lind;
        ind1 >>= 2;
    }
    bool iSA2good = true;
    if (ix.genomeSAindexStart[Lind - 1] + ind1 + 1 < ix.genomeSAindexStart[Lind]) {
        iSA2 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1 + 1);
        nSai++;
        if ((iSA2 & ix.SAiMarkAbsentMaskC) == 0) iSA2 = (iSA2 & ix.SAiMarkNmask) - 1;
        else { iSA2 = ix.nSA - 1; iSA2good = false; }
    } else {
        iSA2 = ix.nSA - 1;
        iSA2good = false;
    }
    const bool iSA1noN = (iSA1 & ix.SAiMarkNmaskC) == 0;
    u32 maxL;
    if (Lind < ix.gSAindexNbases && iSA1noN && iSA2good) {            // the prefix is not in the genome: its longest present part is the answer
        indStartEnd[0] = iSA1; indStartEnd[1] = iSA2;
        Nrep = iSA2 - iSA1 + 1;
        maxL = Lind;
    } else if (plain && iSA1noN && iSA2good && Lind == ix.gSAindexNbases && iSA1 <= iSA2) {
        // a plain SAindex interval: every row starts with the piece's first Lind bases -> keys
        const u32 Lk = Lind;
        const u32 m = pieceLength - Lk < SK_KEY_BASES ? pieceLength - Lk : SK_KEY_BASES;
        const u32 rk = m ? (rk28 >> (2 * (SK_KEY_BASES - m))) << (2 * (SK_KEY_BASES - m)) : 0;   // its first m bases
        u64 b1, b2;
        const u32 kl = m ? keyedWindow(g, keys, scanMax, iSA1, iSA2, rk, m, b1, b2, probes) : 0;
        if (m == 0) { b1 = iSA1; b2 = iSA2; }
        if (kl == m && Lk + m < pieceLength) {                         // rows b1..b2 match all 28 bases and the piece goes on: genome
            if (b1 == b2) { maxL = grpLcpRow(g, ix, R, pieceStart, pieceLength, Lk + m, b1, dirR, bases); probes++; indStartEnd[0] = indStartEnd[1] = b1; }
            else {
                const BlockRes r = groupMaxMappable(g, ix, R, pieceStart, pieceLength, b1, b2, dirR, Lk + m);
                maxL = r.L; indStartEnd[0] = r.b1; indStartEnd[1] = r.b2; probes += r.probes; bases += r.bases;
            }
        } else {
            maxL = Lk + kl;
            indStartEnd[0] = b1; indStartEnd[1] = b2;
        }
        Nrep = indStartEnd[1] - indStartEnd[0] + 1;
    } else {                                                           // interval with N inside the prefix / without an upper bound: search all of it
        const u32 L0 = (iSA2good && iSA1noN) ? Lind : 0;
        const BlockRes r = groupMaxMappable(g, ix, R, pieceStart, pieceLength, iSA1 & ix.SAiMarkNmask, iSA2, dirR, L0);
        maxL = r.L; indStartEnd[0] = r.b1; indStartEnd[1] = r.b2; probes += r.probes; bases += r.bases;
        Nrep = r.b2 - r.b1 + 1;
    }
    maxLout = maxL;
    SAstart = indStartEnd[0];
}

// qualitySplit (SequenceFuns.cpp:411-444) + the chain grid of mapOneRead (ReadAlign_mapOneRead.cpp:37-60): one lane per read, items appended
// with one atomic per warp.  The sort key of an item is the SAindex L-mer of the chain's first search.
__global__ void __launch_bounds__(128) seed_chains_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                          ReadInfo* __restrict__ info, u32 nReads, const __grid_constant__ KeyedArgs ka) {
    const u32 lane = threadIdx.x & 31;
#pragma unroll 1
    for (u32 base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < nReads; base += gridDim.x * blockDim.x) {
        const u32 i = base + lane;
        u32 splitStart[10], splitLen[10], splitFrag[10], nStartOf[10];
        u32 Nsplit = 0, nItems = 0, flags = 0;
        const u8* R = reads + (u64)i * stride;
        if (i < nReads) {
            const u32 Lread = info[i].Lread;
            u32 iR = 0, iS = 0, LgoodMin = 0, iFrag = 0;
            const u32 maxNsplit = (u32)(P.maxNsplit < 10 ? P.maxNsplit : 10);
            while ((iR < Lread) & (iS < maxNsplit)) {
                while (iR < Lread && R[iR] > 3) { if (R[iR] == STAR_MARK_FRAG_SPACER_BASE) iFrag++; iR++; }
                if (iR == Lread) break;
                const u32 iR1 = iR;
                while (iR < Lread && R[iR] <= 3) iR++;
                if ((iR - iR1) > LgoodMin) LgoodMin = iR - iR1;
                if ((iR - iR1) < P.seedSplitMin) continue;
                splitStart[iS] = iR1; splitLen[iS] = iR - iR1; splitFrag[iS] = iFrag;
                iS++;
            }
            Nsplit = iS;
            info[i].Nsplit = (u16)iS;
            info[i].split1_0 = (u16)(iS == 0 ? LgoodMin : splitLen[0]);
            const u64 a = P.seedSearchStartLmax;
            const u64 b = (u64)(P.seedSearchStartLmaxOverLread * (double)(Lread - 1));
            const u64 sLmax = a < b ? a : b;
            for (u32 ip = 0; ip < Nsplit; ip++) {
                const u64 pl = splitLen[ip];
                const u64 Nstart = (P.seedSearchStartLmax > 0 && sLmax < pl) ? pl / sLmax + 1 : 1;
                if (Nstart > 127) { flags = 1; break; }                 // chain ids hold 7 bits of start: such a read takes the tier path
                nStartOf[ip] = (u32)Nstart;
                nItems += 2 * (u32)Nstart - 1;                          // (the reverse chain of start 0 rides on the forward one: flagDirMap)
            }
            if (flags) nItems = 0;
            ka.recCount[i] = 0;
            info[i].cCompare = 0; info[i].cBases = 0;
        }
        // exclusive scan of nItems over the warp, one atomic for the warp's total
        u32 incl = nItems;
        for (u32 o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        u32 wbase = 0;
        if (lane == 0 && total) wbase = atomicAdd(ka.itemCount, total);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (i < nReads) {
            u32 o = wbase + incl - nItems;
            if (nItems && o + nItems > ka.maxItems) {                             // item list full: the read takes the tier path, its slots stay unused
                for (u32 q = o; q < o + nItems && q < ka.maxItems; q++) { ka.items[q].read = 0xffffffffu; ka.itemKey[q] = 0xffffffffu; ka.itemIdx[q] = q; }
                flags = 1; nItems = 0;
            }
            if (flags) info[i].flags |= 1u;
            if (nItems)
                for (u32 ip = 0; ip < Nsplit; ip++) {
                    const u32 pl = splitLen[ip], ps = splitStart[ip], Nstart = nStartOf[ip], Lstart = pl / Nstart;
                    for (u32 iDir = 0; iDir < 2; iDir++)
                        for (u32 istart = iDir; istart < Nstart; istart++) {   // (direction 1 from start 1 on)
                            ChainItem it;
                            it.read = i; it.ps = (u16)ps; it.pl = (u16)pl; it.chainId = (u16)((ip << 8) | (iDir << 7) | istart);
                            it.iFrag = (u8)splitFrag[ip]; it.nStart = (u8)Nstart;
                            // L-mer of the first search of the chain (shorter pieces: left-aligned, so that the order of the keys is the SAindex order)
                            const u64 S = iDir == 0 ? ps + istart * Lstart : ps + pl - istart * Lstart - 1;
                            const u32 len = pl - istart * Lstart;
                            const u32 Lm = ix.gSAindexNbases < len ? ix.gSAindexNbases : len;
                            u32 key = 0;
                            for (u32 ii = 0; ii < Lm; ii++) key = (key << 2) + pieceBase(R, S, iDir == 0, ii);
                            key <<= 2 * (ix.gSAindexNbases - Lm);
                            ka.items[o] = it;
                            ka.itemKey[o] = key;
                            ka.itemIdx[o] = o;
                            o++;
                        }
                }
        }
    }
}

// One group per chain item: the while loop of mapOneRead for (piece, direction, start), the reverse chain of start 0 when flagDirMap stays
// set, and the fixed-length search of --seedSearchLmax.  order: item permutation (sorted by L-mer) or nullptr.
template <u32 GN, int MINB>
__global__ void __launch_bounds__(128, MINB) seed_keyed_search_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                                ReadInfo* __restrict__ info, const u32* __restrict__ order, const __grid_constant__ KeyedArgs ka) {
    typedef GrpT<GN> Grp;
    const Grp g;
    const u32 nItems = *ka.itemCount < ka.maxItems ? *ka.itemCount : ka.maxItems;
    const u32 nGroups = gridDim.x * blockDim.x / Grp::G;
    // The groups of a warp take neighbouring items and meet again before every item: between two meetings each group follows its own
    // chain (different numbers of searches, different paths), and without the meeting they would stay apart for the rest of the kernel —
    // measured: 8.0 instead of 12.6 lanes per issued instruction and 1.6 x the time.
#pragma unroll 1
    for (u32 tb = ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u) / Grp::G; tb < nItems; tb += nGroups) {
        __syncwarp();
        const u32 t = tb + (threadIdx.x & 31u) / Grp::G;
        if (t >= nItems) continue;
        const ChainItem it = ka.items[order ? order[t] : t];
        if (it.read == 0xffffffffu) continue;
        const u8* R = reads + (u64)it.read * stride;
        const u32 ps = it.ps, pl = it.pl, Nstart = it.nStart, Lstart = pl / Nstart;
        const u32 istart = it.chainId & 127u;
        u32 iDir = (it.chainId >> 7) & 1u;
        SeedRec* slab = ka.recs + (u64)it.read * ka.maxRec;
        u32 probes = 0, bases = 0;
        bool flagDirMap = true;
#pragma unroll 1
        for (;;) {                                                          // this chain, then (start 0 only) the reverse chain of the same start
            const u16 chainId = (u16)((it.chainId & ~0x80u) | (iDir << 7));
            u32 Lmapped = 0, k = 0;
            if (flagDirMap || istart > 0) {
#pragma unroll 1
                while (istart * Lstart + Lmapped + P.seedMapMin < pl) {
                    const u64 Shift = iDir == 0 ? (u64)(ps + istart * Lstart + Lmapped) : (u64)(ps + pl - istart * Lstart - 1 - Lmapped);
                    const u32 seedLength = pl - Lmapped - istart * Lstart;
                    u32 L, nSai;
                    u64 Nrep, SAstart;
                    groupSearch(g, ix, ka.saKeys, ka.scanMax, R, Shift, seedLength, iDir == 0, true, L, Nrep, SAstart, nSai, probes, bases);
                    if (iDir == 0 && istart == 0 && Lmapped == 0 && Shift + L == pl) flagDirMap = false;
                    if (g.lane == 0) {
                        const u32 slot = atomicAdd(ka.recCount + it.read, 1u);
                        if (slot < ka.maxRec) {
                            SeedRec r;
                            r.SAstart = SAstart; r.Nrep = Nrep > 0xffffffffULL ? 0xffffffffu : (u32)Nrep; r.Shift = (u16)Shift; r.L = (u16)L;
                            r.chainId = chainId; r.k = (u8)(k < 254 ? k : 254); r.nSai = (u8)nSai; r.iFrag = it.iFrag; r.pad_[0] = r.pad_[1] = r.pad_[2] = 0;
                            slab[slot] = r;
                        }
                    }
                    k++;
                    if (__builtin_expect(k >= 254, 0)) {   // more searches in one chain than a record can number: the read takes the tier path
                        if (g.lane == 0) atomicAdd(ka.recCount + it.read, ka.maxRec + 1u);
                        break;
                    }
                    Lmapped += L;
                }
            }
            if (__builtin_expect(P.seedSearchLmax > 0, 0)) {                // ReadAlign_mapOneRead.cpp:81-87
                const u64 Shift = iDir == 0 ? (u64)(ps + istart * Lstart) : (u64)(ps + pl - istart * Lstart - 1);
                const u64 room = iDir == 0 ? (ps + pl - Shift) : (Shift + 1);
                u32 L, nSai;
                u64 Nrep, SAstart;
                const u32 sl = (u32)(P.seedSearchLmax < room ? P.seedSearchLmax : room);
                groupSearch(g, ix, ka.saKeys, ka.scanMax, R, Shift, sl, iDir == 0, iDir == 0 || sl <= Shift + 1 - ps, L, Nrep, SAstart, nSai, probes, bases);
                if (g.lane == 0) {
                    const u32 slot = atomicAdd(ka.recCount + it.read, 1u);
                    if (slot < ka.maxRec) {
                        SeedRec r;
                        r.SAstart = SAstart; r.Nrep = Nrep > 0xffffffffULL ? 0xffffffffu : (u32)Nrep; r.Shift = (u16)Shift; r.L = (u16)L;
                        r.chainId = chainId; r.k = 255; r.nSai = (u8)nSai; r.iFrag = it.iFrag; r.pad_[0] = r.pad_[1] = r.pad_[2] = 0;
                        slab[slot] = r;
                    }
                }
            }
            if (istart == 0 && iDir == 0) { iDir = 1; continue; }           // the reverse chain of start 0 depends on flagDirMap: same group
            break;
        }
        if (g.lane == 0) { atomicAdd(&info[it.read].cCompare, probes); atomicAdd(&info[it.read].cBases, bases); }
    }
}

}  // namespace starb
