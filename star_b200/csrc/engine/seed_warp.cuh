// seed_warp.cuh — maximal-mappable-prefix seed search of ONE read by ONE warp (all 32 lanes, warp-uniform control flow).
//
// Same results as seed_search_kernel (seed.cu) / the reference (ReadAlign_mapOneRead.cpp:37-93, ReadAlign_maxMappableLength2strands.cpp:5-115,
// SuffixArrayFuns.cpp:10-207, ReadAlign_storeAligns.cpp:10-160) with a different search strategy inside maxMappableLength:
// the answer — the maximal match length L over the SA rows of the start interval and the block of rows attaining it — does not
// depend on the probing order, so the warp narrows the interval with 32 probes per step until it has at most 32 rows and then
// examines the whole window in ONE step (measured on the CPU emulation of the test infrastructure: 1.30 dependent steps per search
// instead of the binary search's 7.48; identical results on 4.5 M searches).  The SA words of a window are contiguous (coalesced);
// every lane compares the read with its own suffix.
//
// Written against the warp interface of warp_prims.cuh: compiled for the GPU by seed.cu (DevWarp) and for the host by
// oracle/warp_emul.cpp (HostWarp: 32 threads), where tests compare every stored piece with the oracle's.
#pragma once
#include "dev.cuh"

namespace starb {

struct SeedWarpOut {
    Piece* PC;                 // slab of this read (global memory / host array)
    u32 maxP;
    u32 nP, nA, multNmin, multNminL, flags;
    u32 Nsplit, split1_0;
    u32 searches, saiWords, probes, rounds;   // work counters (probes / rounds of THIS search strategy, not the reference's compare calls)
    u32 basesLane;             // genome bases examined by THIS lane (the only per-lane value; summed over the warp by the caller)
};

// 8 bytes starting at an arbitrary address as one little-endian word (byte k = p[k]): two aligned 64-bit loads and a funnel shift.
// Up to 7 bytes before / 15 after p are touched: the genome carries 256 bytes of padding, the read rows 16 bytes of slack on both sides.
SB_DEV u64 load8generic(const u8* p) {
    const uintptr_t ad = (uintptr_t)p;
    const u64* a = (const u64*)(ad & ~(uintptr_t)7);
    const u32 sh = (u32)(ad & 7) * 8;
    const u64 lo = a[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (a[1] << (64 - sh));
}
SB_DEV u64 load8global(const u8* p) {   // read-only data in global memory
    const uintptr_t ad = (uintptr_t)p;
    const u64* a = (const u64*)(ad & ~(uintptr_t)7);
    const u32 sh = (u32)(ad & 7) * 8;
    const u64 lo = SB_LDG(a);
    if (sh == 0) return lo;
    return (lo >> sh) | (SB_LDG(a + 1) << (64 - sh));
}
SB_DEV u64 bswap64(u64 v) {
#ifdef STAR_WARP_HOST_EMUL
    return __builtin_bswap64(v);
#else
    const u32 lo = (u32)v, hi = (u32)(v >> 32);
    return ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
#endif
}

// match length of the read piece against the suffix of SA row iSA, starting at offset L (all rows of the current window share the first L
// bases); the four cases of compareSeqToGenome (SuffixArrayFuns.cpp:10-104) in one loop with per-lane direction flags, 8 bases per step:
// the 8 read bases and the 8 genome bases are gathered into two words (byte k = base ii+k in comparison order), XOR-ed, and the first
// non-zero byte is the first mismatch.  Returns the length (N = full match) and compRes (read > suffix in SA order).
template <class W>
SB_DEV u32 lcpRow(const DevIndex& ix, const u8* R, u64 S, u32 N, u32 L, u64 iSA, bool dirR, bool& compRes, u32& bases) {
    u64 SAstr = packedGet(ix.SA, ix.saBits, iSA);
    const bool dirG = (SAstr >> ix.GstrandBit) == 0;
    SAstr &= ix.GstrandMask;
    const bool compl_ = dirR != dirG;                 // the read is compared as its complement (piece bases are always 0..3: 3-x == x^3)
    const u8* g0 = ix.G + (dirG ? (long long)(SAstr + L) : (long long)(ix.nGenome - 1 - SAstr - L));   // base ii is g0[+ii] / g0[-ii]
    const u8* r0 = R + (dirR ? (long long)(S + L) : (long long)S - (long long)L);                      // base ii is r0[+ii] / r0[-ii]
    const u32 n = N - L;
    for (u32 ii = 0; ii < n; ii += 8) {
        u64 rs = dirR ? load8generic(r0 + ii) : bswap64(load8generic(r0 - (long long)ii - 7));
        const u64 gs = dirG ? load8global(g0 + ii) : bswap64(load8global(g0 - (long long)ii - 7));
        if (compl_) rs ^= 0x0303030303030303ULL;
        const u64 x = rs ^ gs;
        if (x) {
            const u32 k = (u32)SB_CTZ64(x) >> 3;       // first differing base of this step
            if (ii + k < n) {
                const u8 sv = (u8)(rs >> (8 * k)), gv = (u8)(gs >> (8 * k));
                compRes = dirG ? (sv > gv) : !(sv > gv || gv > 3);
                bases += ii + k + 1;
                return ii + k + L;
            }
            break;                                     // the difference lies beyond the piece
        }
    }
    compRes = false;
    bases += n;
    return N;
}

// The block [b1,b2] of SA rows (clamped to [lo,hi]) whose match length with the piece is maximal, and that length.
template <class W>
SB_DEV u64 warpMaxMappableLength(const W& w, const DevIndex& ix, const u8* R, u64 S, u32 N, u64 lo, u64 hi, bool dirR, u32& L, u64* indStartEnd,
                                 SeedWarpOut& o) {
    const u32 lane = w.lane;
    u64 i1 = lo, i2 = hi, i3 = lo;
    u32 L3 = 0, Lc = L;
    bool have = false;
    // (1) narrowing: 32 probes spread over the window (both ends included) until the window has at most 32 rows
    while (i2 - i1 + 1 > 32) {
        const u64 row = i1 + (u64)(((unsigned __int128)(i2 - i1) * lane) / 31);
        bool c;
        const u32 Lj = lcpRow<W>(ix, R, S, N, Lc, row, dirR, c, o.basesLane);
        o.rounds++; o.probes += 32;
        const u32 fullMask = w.ballot(Lj == N);
        const u32 gtMask = w.ballot(Lj != N && c);                 // read > suffix: the insertion point is to the right of this probe
        if (fullMask) { const int src = SB_FFS(fullMask) - 1; i3 = w.shfl64(row, src); L3 = N; have = true; break; }
        if (!gtMask) { i3 = w.shfl64(row, 0); L3 = w.shfl(Lj, 0); have = true; break; }              // the read sorts before the first row
        const int jLast = 31 - SB_CLZ(gtMask);
        if (jLast == 31) { i3 = w.shfl64(row, 31); L3 = w.shfl(Lj, 31); have = true; break; }        // ... after the last row
        const u64 n1 = w.shfl64(row, jLast), n2 = w.shfl64(row, jLast + 1);
        const u32 l1 = w.shfl(Lj, jLast), l2 = w.shfl(Lj, jLast + 1);
        i1 = n1; i2 = n2;
        Lc = l1 < l2 ? l1 : l2;
    }
    u64 b1, b2;
    bool seenLeft = false, seenRight = false;
    if (!have) {
        // (2) one step over every row of the window
        const u32 rows = (u32)(i2 - i1 + 1);
        const bool valid = lane < rows;
        bool c = false;
        const u32 Lj = valid ? lcpRow<W>(ix, R, S, N, Lc, i1 + lane, dirR, c, o.basesLane) : 0;
        o.rounds++; o.probes += rows;
        L3 = (u32)w.reduceMax(valid ? (int)Lj : -1);
        const u32 eq = w.ballot(valid && Lj == L3);
        const int jm = SB_FFS(eq) - 1;                 // first row with the maximal length; its block = the run of set bits around it
        u32 run = eq >> jm;                                        // bit 0 = jm
        const int len = run == 0xffffffffu >> jm ? 32 - jm : SB_CTZ(~run);   // rows of the run starting at jm (rows below jm are shorter: jm is the first)
        const int j1 = jm, j2 = jm + len - 1;
        b1 = i1 + (u64)j1; b2 = i1 + (u64)j2;
        i3 = b1;
        seenLeft = j1 > 0; seenRight = (u32)j2 < rows - 1;
    } else {
        b1 = b2 = i3;
    }
    // (3) the block may continue beyond the rows seen so far (repeats): 32-ary boundary searches, clamped to [lo, hi]
    for (int side = 0; side < 2; side++) {
        const bool left = side == 0;
        if (left ? (seenLeft || b1 <= lo) : (seenRight || b2 >= hi)) continue;
        const u64 inRow = left ? b1 : b2, outRow = left ? lo : hi;
        bool c;
        const u32 Lout = lcpRow<W>(ix, R, S, L3, L, outRow, dirR, c, o.basesLane);      // (same row on every lane: one broadcast load)
        o.rounds++; o.probes += 1;
        u64 res;
        if (Lout >= L3) res = outRow;
        else {
            u64 a = outRow, b = inRow;      // a: shorter match, b: match >= L3
            u32 La = Lout;
            while (left ? a + 1 < b : b + 1 < a) {
                const u64 span = (left ? b - a : a - b) - 1;
                const u32 np = span < 32 ? (u32)span : 32;
                const bool valid = lane < np;
                const u64 step = np == span ? 1 + lane : (u64)(((unsigned __int128)(span + 1) * (lane + 1)) / (np + 1));
                const u64 row = left ? a + step : a - step;
                const u32 Lj = valid ? lcpRow<W>(ix, R, S, L3, La, row, dirR, c, o.basesLane) : 0;
                o.rounds++; o.probes += np;
                const u32 inMask = w.ballot(valid && Lj >= L3);
                if (!inMask) { a = w.shfl64(row, (int)np - 1); La = w.shfl(Lj, (int)np - 1); }
                else {
                    const int jIn = SB_FFS(inMask) - 1;
                    b = w.shfl64(row, jIn);
                    if (jIn > 0) { a = w.shfl64(row, jIn - 1); La = w.shfl(Lj, jIn - 1); }
                }
            }
            res = b;
        }
        if (left) b1 = res; else b2 = res;
    }
    L = L3;
    indStartEnd[0] = b1; indStartEnd[1] = b2;
    return b2 - b1 + 1;
}

// ReadAlign_storeAligns.cpp:10-51,140-160 (OPTIM_STOREaligns_SIMPLE), warp-uniform: every lane keeps the same counters, lane 0 edits the slab
template <class W>
SB_DEV void warpStoreAligns(const W& w, SeedWarpOut& st, const star_params_t& P, u32 iDir, u64 Shift, u64 Nrep, u64 L, u64 SAstart, u32 iFrag) {
    if (Nrep > P.seedMultimapNmax) {
        if (Nrep < st.multNmin || st.multNmin == 0) { st.multNmin = (u32)Nrep; st.multNminL = (u32)L; }
        return;
    }
    st.nA += (u32)Nrep;
    const u32 rStart = (u32)(iDir == 0 ? Shift : Shift + 1 - L);
    int iP;
    for (iP = (int)st.nP - 1; iP >= 0; iP--) {
        const u32 r0 = st.PC[iP].rStart;
        if (r0 <= rStart) {
            if (r0 == rStart && st.PC[iP].Length < L) continue;
            if (r0 == rStart && st.PC[iP].Length == L) return;
            break;
        }
    }
    iP = iP + 1;
    if (st.nP + 1 > P.seedPerReadNmax) { st.flags |= 2; return; }   // fatal in the reference (:46-51)
    if (st.nP + 1 > st.maxP) { st.flags |= 1; return; }             // slab of this tier full: the read is redone with a bigger one
    w.sync();
    if (w.lane == 0) {
        for (int ii = (int)st.nP - 1; ii >= iP; ii--) st.PC[ii + 1] = st.PC[ii];
        Piece p;
        p.SAstart = SAstart; p.rStart = (u16)rStart; p.Length = (u16)L; p.Nrep = (u16)Nrep; p.Dir = (u8)iDir; p.iFrag = (u8)iFrag;
        st.PC[iP] = p;
    }
    w.sync();
    st.nP++;
    if (Nrep != 1) {
        if (Nrep < st.multNmin || st.multNmin == 0) { st.multNmin = (u32)Nrep; st.multNminL = (u32)L; }
    }
}

// ReadAlign_maxMappableLength2strands.cpp:5-115 with gSAsparseD == 1
template <class W>
SB_DEV void warpMaxMappableLength2strands(const W& w, const DevIndex& ix, const star_params_t& P, const u8* R, SeedWarpOut& st, u64 pieceStart, u64 pieceLength,
                                          u32 iDir, u64& maxLbest, u32 iFrag) {
    u64 Nrep = 0, indStartEnd[2] = {0, 0};
    u32 maxL = 0;
    const bool dirR = iDir == 0;
    st.searches++;
    const u64 Lmax = ix.gSAindexNbases < pieceLength ? ix.gSAindexNbases : pieceLength;
    u64 ind1 = 0;
    if (dirR) { for (u64 ii = 0; ii < Lmax; ii++) { ind1 <<= 2; ind1 += (u64)R[pieceStart + ii]; } }
    else { for (u64 ii = 0; ii < Lmax; ii++) { ind1 <<= 2; ind1 += (3 - (u64)R[pieceStart - ii]); } }
    u64 Lind = Lmax;
    u64 iSA1 = 0, iSA2 = 0;
    while (Lind > 0) {
        iSA1 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1);
        st.saiWords++;
        if ((iSA1 & ix.SAiMarkAbsentMaskC) == 0) break;
        --Lind;
        ind1 = ind1 >> 2;
    }
    bool iSA2good = true;
    if (ix.genomeSAindexStart[Lind - 1] + ind1 + 1 < ix.genomeSAindexStart[Lind]) {
        iSA2 = packedGet(ix.SAi, ix.saiBits, ix.genomeSAindexStart[Lind - 1] + ind1 + 1);
        st.saiWords++;
        if ((iSA2 & ix.SAiMarkAbsentMaskC) == 0) iSA2 = (iSA2 & ix.SAiMarkNmask) - 1;
        else { iSA2 = ix.nSA - 1; iSA2good = false; }
    } else {
        iSA2 = ix.nSA - 1;
        iSA2good = false;
    }
    const bool iSA1noN = (iSA1 & ix.SAiMarkNmaskC) == 0;
    if (Lind < ix.gSAindexNbases && iSA1noN && iSA2good) {
        indStartEnd[0] = iSA1; indStartEnd[1] = iSA2;
        Nrep = iSA2 - iSA1 + 1;
        maxL = (u32)Lind;
    } else if (iSA1 == iSA2 && iSA1noN && iSA2good) {
        indStartEnd[0] = indStartEnd[1] = iSA1;
        Nrep = 1;
        bool c;
        maxL = lcpRow<W>(ix, R, pieceStart, (u32)pieceLength, (u32)Lind, iSA1, dirR, c, st.basesLane);   // same row on every lane
        st.rounds++; st.probes++;
    } else {
        maxL = (iSA2good && iSA1noN) ? (u32)Lind : 0;
        Nrep = warpMaxMappableLength<W>(w, ix, R, pieceStart, (u32)pieceLength, iSA1 & ix.SAiMarkNmask, iSA2, dirR, maxL, indStartEnd, st);
    }
    maxLbest = maxL;
    warpStoreAligns<W>(w, st, P, iDir, pieceStart, Nrep, maxL, indStartEnd[0], iFrag);
}

// One read: qualitySplit (SequenceFuns.cpp:411-444) + the search schedule of ReadAlign::mapOneRead (ReadAlign_mapOneRead.cpp:37-93).
// R = Read1[0] of the read (mate1 | spacer | revcomp(mate2), codes 0..4 / spacer), readable by every lane.
template <class W>
SB_DEV void warpSeedRead(const W& w, const DevIndex& ix, const star_params_t& P, const u8* R, u32 Lread, SeedWarpOut& st) {
    st.nP = 0; st.nA = 0; st.multNmin = 0; st.multNminL = 0; st.flags = 0;
    st.searches = 0; st.saiWords = 0; st.probes = 0; st.rounds = 0; st.basesLane = 0;
    u32 splitStart[10], splitLen[10], splitFrag[10];
    u32 Nsplit = 0;
    {
        u32 iR = 0, iS = 0, LgoodMin = 0, iFrag = 0;
        const u32 maxNsplit = (u32)(P.maxNsplit < 10 ? P.maxNsplit : 10);
        while ((iR < Lread) & (iS < maxNsplit)) {
            while (iR < Lread && R[iR] > 3) {
                if (R[iR] == STAR_MARK_FRAG_SPACER_BASE) iFrag++;
                iR++;
            }
            if (iR == Lread) break;
            const u32 iR1 = iR;
            while (iR < Lread && R[iR] <= 3) iR++;
            if ((iR - iR1) > LgoodMin) LgoodMin = iR - iR1;
            if ((iR - iR1) < P.seedSplitMin) continue;
            splitStart[iS] = iR1; splitLen[iS] = iR - iR1; splitFrag[iS] = iFrag;
            iS++;
        }
        Nsplit = iS;
        st.Nsplit = iS;
        st.split1_0 = iS == 0 ? LgoodMin : splitLen[0];
    }
    const u64 a = P.seedSearchStartLmax;
    const u64 b = (u64)(P.seedSearchStartLmaxOverLread * (double)(Lread - 1));
    const u64 seedSearchStartLmax = a < b ? a : b;
    for (u32 ip = 0; ip < Nsplit && !st.flags; ip++) {
        const u64 pl = splitLen[ip], ps = splitStart[ip];
        const u64 Nstart = (P.seedSearchStartLmax > 0 && seedSearchStartLmax < pl) ? pl / seedSearchStartLmax + 1 : 1;
        const u64 Lstart = pl / Nstart;
        bool flagDirMap = true;
        for (u32 iDir = 0; iDir < 2; iDir++) {
            for (u64 istart = 0; istart < Nstart; istart++) {
                if (flagDirMap || istart > 0) {
                    u64 Lmapped = 0;
                    while (istart * Lstart + Lmapped + P.seedMapMin < pl) {
                        const u64 Shift = iDir == 0 ? (ps + istart * Lstart + Lmapped) : (ps + pl - istart * Lstart - 1 - Lmapped);
                        const u64 seedLength = pl - Lmapped - istart * Lstart;
                        u64 L;
                        warpMaxMappableLength2strands<W>(w, ix, P, R, st, Shift, seedLength, iDir, L, splitFrag[ip]);
                        if (iDir == 0 && istart == 0 && Lmapped == 0 && Shift + L == pl) flagDirMap = false;
                        Lmapped += L;
                        if (st.flags) break;
                    }
                }
                if (P.seedSearchLmax > 0 && !st.flags) {   // ReadAlign_mapOneRead.cpp:81-87: fixed-length search from every start (off by default)
                    const u64 Shift = iDir == 0 ? (ps + istart * Lstart) : (ps + pl - istart * Lstart - 1);
                    const u64 room = iDir == 0 ? (ps + pl - Shift) : (Shift + 1);
                    u64 L;
                    warpMaxMappableLength2strands<W>(w, ix, P, R, st, Shift, P.seedSearchLmax < room ? P.seedSearchLmax : room, iDir, L, splitFrag[ip]);
                }
                if (st.flags) break;
            }
            if (st.flags) break;
        }
    }
}

}  // namespace starb
