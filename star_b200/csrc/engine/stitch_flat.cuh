// stitch_flat.cuh — heavy reads, flattened over the whole chunk (included at the end of stitch.cu, namespace starb).
//
// stitch_heavy_kernel gives one warp to one read and keeps the read's three phases inside that warp: windows (cooperative),
// sub-tree evaluation (lanes pull tasks), ordered recording (lane 0).  Measured on B200 (profiles/r01_summary.md): the lanes of a
// warp wait for the longest sub-tree of THEIR read and for lane 0's recording, 5.7 of 32 lanes execute on average, and the
// per-lane DFS state in local memory (interleaved by lane) is cached with a handful of useful bytes per 32-byte sector.
// Here the three phases are three kernels over ALL heavy reads of the chunk, so that no lane ever waits for a read:
//
//   flat_setup_kernel   one warp per read: window creation / seed assignment (same cooperative code as the heavy kernel), then
//                       the read is exported to the flat pool in HBM: R0 | R2 | FlatWin[nWin] | Seed[nSeeds], plus one FlatTask
//                       per prefix sub-tree of every window (task ids of a read are contiguous and ascending = the reference's
//                       DFS order, windows in window order);
//   flat_dfs_warp_kernel   one WARP per task (all lanes execute the same scalar path), global ticket: the pure part of the recursion
//                       (stitch chain, end extension, filters, score) for every leaf of the sub-tree; surviving leaves become
//                       16-byte candidates in the task's FlatOut (first candidate inline, more in 128-byte blocks);
//   flat_record_warp_kernel  one WARP per read: the reference's order-dependent part (maxScoreMate, record test, blocksOverlap dedup,
//                       ordered insert, multMapSelect) over the candidates in window / task / leaf order.
//
// Results are identical to the sequential recursion for the same reason as in the heavy kernel (evalLeaf is a pure function of
// the path; recordLeaf sees the leaves in DFS order).

__device__ __forceinline__ u32 flatReadStride(u32 Lread) { return (Lread + 16) & ~15u; }

template <int MINB>
__global__ void __launch_bounds__(128, MINB) flat_setup_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, const u8* __restrict__ reads, u32 stride,
                                                   ReadInfo* __restrict__ info, const Piece* __restrict__ pieces, u32 nHeavy, const u32* __restrict__ heavyList,
                                                   const u64* __restrict__ heavyOff, const u8* __restrict__ heavyPool, u32* __restrict__ counter,
                                                   u8* __restrict__ arenas, Caps caps, star_read_result_t* __restrict__ results,
                                                   star_align_t* __restrict__ staged, u32 smemStride, FlatArgs fa, u32 kBase) {
    extern __shared__ u8 smem[];
    const u32 lane = threadIdx.x & 31;
    const u32 warpInBlock = threadIdx.x >> 5;
    const u32 warpsPerBlock = blockDim.x >> 5;
    const u32 gwarp = blockIdx.x * warpsPerBlock + warpInBlock;
    const u32 perWarp = (2 * smemStride + 32 + caps.maxW * (u32)sizeof(Window) + (caps.maxW + 4) * 4 + ((caps.maxW + 3) & ~3u) + 15) & ~15u;
    u8* R0 = smem + (size_t)warpInBlock * perWarp;
    u8* R2 = R0 + smemStride;
    u32* sh = (u32*)(R0 + 2 * smemStride);
    Window* swin = (Window*)(R0 + 2 * smemStride + 32);
    u32* taskStart = (u32*)(swin + caps.maxW);
    u8* depthOf = (u8*)(taskStart + caps.maxW + 4);
    Lane ln;
    ln.cur = nullptr; ln.leaf = nullptr; ln.stack = nullptr; ln.ph = nullptr;   // this kernel never stitches
    ln.ix = &ix; ln.P = &P; ln.R0 = R0; ln.R2 = R2; ln.R = R0; ln.caps = caps;
    {
        u8* a = arenas + (u64)gwarp * caps.arenaBytes;
        ln.win = (Window*)a; a += (u64)caps.maxW * sizeof(Window);
        ln.wa = (Seed*)a; a += (u64)caps.maxW * caps.spw * sizeof(Seed);
        ln.pool = (DevTr*)a; a += (u64)caps.maxTr * sizeof(DevTr);
        a += 2 * sizeof(DevTr) + (u64)(caps.spw + 2) * sizeof(Frame);
        ln.trPtr = (u16*)a; a += (u64)caps.maxTr * sizeof(u16);
        ln.winBase = (u16*)a; a += (u64)caps.maxW * sizeof(u16);
        ln.winN = (u16*)a;
    }
    ln.memo = nullptr; ln.memoMask = 0; ln.memoBase = 0; ln.memoHit = 0; ln.memoMiss = 0; ln.lastSeed = -1; ln.coop = 0;
    WarpWin ww;
    ww.swin = swin; ww.wa = ln.wa; ww.spw = caps.spw; ww.lane = lane; ww.nW = 0;
    long long tSetup = 0;
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
        FlatRec rec;
        rec.poolOff = 0; rec.read = i; rec.nWin = 0; rec.taskBase = 0; rec.nTasks = 0; rec.over = 0; rec.done = 0; rec.saEnum = 0;
        rec.Lread = Lread; rec.mmMax = ri.outFilterMismatchNmaxTotal; rec.readLength[0] = ri.readLength[0]; rec.readLength[1] = ri.readLength[1];
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
                rec.done = 1;
                fa.recs[kBase + k] = rec;
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
        __syncwarp();
        u32 nWin = 0;
        u32 overReason = 0;
        // (the depth bytes behind the task table are not in use before the windows are final: they hold the bin filter of the assignment phase)
        warpBuildWindows(ln, ww, ix, P, ri, i, fa.slabByPos ? k : i, pieces, heavyOff, heavyPool, caps, swin, lane, nWin, overReason, taskStart,
                         (u32*)depthOf, 31u - (u32)__clz((int)(((caps.maxW + 3) & ~3u) * 8u)));
        __syncwarp();
        // ---- task table (lane 0), allocation in the flat pool / task array
        u32 nTasks = 0, nWinC = 0, nSeeds = 0, taskBase = 0;
        u64 poolOff = 0;
        const u32 rs = flatReadStride(Lread);
        if (lane == 0) {
            if (!overReason) {
                u32 shift = 0;
                #pragma unroll 1
                for (;;) {
                    u32 tot = 0;
                    nWinC = 0; nSeeds = 0;
                    #pragma unroll 1
                    for (u32 w = 0; w < nWin; w++) {
                        u32 a = swin[w].nWA;
                        u32 d = a <= fa.splitMin ? 0 : (a - fa.splitMin > 8 ? 8 : a - fa.splitMin);
                        d = d > shift ? d - shift : 0;
                        taskStart[w] = tot; depthOf[w] = (u8)d;
                        if (a) { tot += 1u << d; nWinC++; nSeeds += a; }
                    }
                    taskStart[nWin] = tot;
                    if (tot <= fa.maxTasksPerRead) { nTasks = tot; break; }
                    shift++;
                }
                const u64 bytes = ((u64)2 * rs + (u64)nWinC * sizeof(FlatWin) + (u64)nSeeds * sizeof(Seed) + 15) & ~15ULL;
                poolOff = atomicAdd(&fa.bumps[0], (unsigned long long)bytes);
                if (poolOff + bytes > fa.poolBytes) overReason = 5;
                if (nTasks) {
                    const u64 tb = atomicAdd(&fa.bumps[1], (unsigned long long)nTasks);
                    if (tb + nTasks > fa.maxTasks) {
                        overReason = 5;
                        #pragma unroll 1
                        for (u64 t = tb; t < tb + nTasks && t < fa.maxTasks; t++) { FlatTask h; h.k = FLAT_NONE; h.w = 0; h.bits = 0; fa.tasks[t] = h; }
                    }
                    taskBase = (u32)tb;
                }
            }
            rec.over = overReason;
            rec.saEnum = (u32)ln.saEnum;
            if (!overReason) { rec.poolOff = poolOff; rec.nWin = nWinC; rec.taskBase = taskBase; rec.nTasks = nTasks; }
            fa.recs[kBase + k] = rec;
        }
        overReason = __shfl_sync(0xffffffffu, overReason, 0);
        poolOff = __shfl_sync(0xffffffffu, poolOff, 0);
        taskBase = __shfl_sync(0xffffffffu, taskBase, 0);
        nWinC = __shfl_sync(0xffffffffu, nWinC, 0);
        if (!overReason) {
            u8* rp = fa.pool + poolOff;
            #pragma unroll 1
            for (u32 b = lane * 4; b < rs; b += 128) {   // both orientations (32-bit copies; rs and smemStride are multiples of 4)
                u32 v0 = 0, v2 = 0;
                if (b + 4 <= smemStride) { v0 = *(const u32*)(R0 + b); v2 = *(const u32*)(R2 + b); }
                *(u32*)(rp + b) = v0;
                *(u32*)(rp + rs + b) = v2;
            }
            FlatWin* fw = (FlatWin*)(rp + 2 * (u64)rs);
            Seed* fs = (Seed*)(fw + nWinC);
            u32 wc = 0, so = 0;
            const u32 kk = kBase + k;
            #pragma unroll 1
            for (u32 w = 0; w < nWin; w++) {   // uniform loop
                const u32 a = swin[w].nWA;
                if (a == 0) continue;
                const u32 d = depthOf[w], ts = taskStart[w];
                if (lane == 0) {
                    FlatWin f; f.Chr = swin[w].Chr; f.nWA = (u16)a; f.Str = swin[w].Str; f.depth = (u8)d; f.seedOff = so; f.taskStart = ts;
                    fw[wc] = f;
                }
                const Seed* src = ln.wa + (u64)w * caps.spw;
                #pragma unroll 1
                for (u32 q = lane; q < a; q += 32) fs[so + q] = src[q];
                #pragma unroll 1
                for (u32 q = lane; q < (1u << d); q += 32) { FlatTask h; h.k = kk; h.w = (u16)wc; h.bits = (u16)q; fa.tasks[(u64)taskBase + ts + q] = h; }
                wc++; so += a;
            }
        }
        __syncwarp();
        tSetup += clock64() - t0;
    }
    PROF_ADD(16, tSetup);
}

// ---- sub-tree task kernel, warp-uniform.  Measured on B200 with one task per LANE (removed; profiles/r01_summary.md): 1.97 lanes
// execute per instruction and 32-task batches with 1.7 working lanes take the same time as 12 working lanes — every lane walks its
// own branchy path and SIMT serialises them.  So a WARP owns one task at a time and all 32 lanes execute the same scalar path (no divergence, every local-memory access is
// one full 128-byte line); the DFS transcript, leaf copy and undo records are ONE copy per warp in shared memory; the byte loops
// (end extension, junction scan, gap mismatches) are done cooperatively, 32 bases per step (coop* functions in stitch.cu).
// Tasks are fetched 32 at a time (lane j prefetches the descriptor, read header and window of task base+j).
#define FLAT_WARP_SMEM (2 * (u32)sizeof(DevTr) + STAR_UNDO_DEPTH * (u32)sizeof(Frame) + 64)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) flat_dfs_warp_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, FlatArgs fa, u32* __restrict__ counter, Caps caps) {
    extern __shared__ u8 smem[];
    const u32 lane = threadIdx.x & 31;
    const u32 warpInBlock = threadIdx.x >> 5;
    u64 nT = fa.bumps[1];
    if (nT > fa.maxTasks) nT = fa.maxTasks;
    u8* ws = smem + (size_t)warpInBlock * FLAT_WARP_SMEM;
    Lane ln;
    ln.stack = (Frame*)ws;
    ln.cur = (DevTr*)(ws + STAR_UNDO_DEPTH * sizeof(Frame));
    ln.leaf = ln.cur + 1;
    ln.ph = (u8*)(ln.leaf + 1);
    ln.ix = &ix; ln.P = &P; ln.R0 = nullptr; ln.R2 = nullptr; ln.R = nullptr; ln.caps = caps;
    ln.win = nullptr; ln.wa = nullptr; ln.pool = nullptr; ln.trPtr = nullptr; ln.winBase = nullptr; ln.winN = nullptr;
    ln.memo = nullptr; ln.memoMask = 0; ln.memoBase = 0; ln.memoHit = 0; ln.memoMiss = 0; ln.lastSeed = -1; ln.coop = 0;
    ln.overflow = 0; ln.saEnum = 0; ln.nodes = 0; ln.leaves = 0; ln.maxScoreMate[0] = ln.maxScoreMate[1] = 0;
    ln.Lread = 0; ln.readLength[0] = ln.readLength[1] = 0; ln.outFilterMismatchNmaxTotal = 0;
    ln.forceDepth = 0; ln.forceBits = 0;
    ln.coop = 1;
    u64 trCur = 0, trEnd = 0;
    long long tStart = clock64();
    #pragma unroll 1
    for (;;) {
        u32 base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((u64)base >= nT) break;
        // lane j prefetches task base+j
        const u64 tMine = (u64)base + lane;
        FlatTask tk; tk.k = FLAT_NONE; tk.w = 0; tk.bits = 0;
        if (tMine < nT) tk = fa.tasks[tMine];
        u64 poolOffM = 0; u32 LreadM = 0, mmM = 0, rlM = 0, nWinM = 0;
        FlatWin WM; WM.Chr = 0; WM.nWA = 0; WM.Str = 0; WM.depth = 0; WM.seedOff = 0; WM.taskStart = 0;
        if (tk.k != FLAT_NONE) {
            const FlatRec* r = &fa.recs[tk.k];
            poolOffM = r->poolOff; LreadM = r->Lread; mmM = r->mmMax; rlM = (u32)r->readLength[0] | ((u32)r->readLength[1] << 16); nWinM = r->nWin;
            WM = ((const FlatWin*)(fa.pool + poolOffM + 2 * (u64)flatReadStride(LreadM)))[tk.w];
        }
        const u32 vmask = __ballot_sync(0xffffffffu, tk.k != FLAT_NONE);
        #pragma unroll 1
        for (u32 j = 0; j < 32; j++) {
            if (!((vmask >> j) & 1u)) continue;
            const u32 k = __shfl_sync(0xffffffffu, tk.k, j);
            const u32 bits = __shfl_sync(0xffffffffu, (u32)tk.bits, j);
            const u64 poolOff = __shfl_sync(0xffffffffu, poolOffM, j);
            const u32 Lread = __shfl_sync(0xffffffffu, LreadM, j);
            const u32 rl = __shfl_sync(0xffffffffu, rlM, j);
            const u32 nWinK = __shfl_sync(0xffffffffu, nWinM, j);
            const u32 Chr = __shfl_sync(0xffffffffu, WM.Chr, j);
            const u32 wpk = __shfl_sync(0xffffffffu, (u32)WM.nWA | ((u32)WM.Str << 16) | ((u32)WM.depth << 24), j);
            const u32 seedOff = __shfl_sync(0xffffffffu, WM.seedOff, j);
            ln.outFilterMismatchNmaxTotal = __shfl_sync(0xffffffffu, mmM, j);
            const u32 nA = wpk & 0xffffu, Str = (wpk >> 16) & 0xffu, depth = wpk >> 24;
            const u64 t = (u64)base + j;
            const u32 rs = flatReadStride(Lread);
            const u8* rp = fa.pool + poolOff;
            const Seed* WA = (const Seed*)(rp + 2 * (u64)rs + (u64)nWinK * sizeof(FlatWin)) + seedOff;
            ln.Lread = Lread; ln.readLength[0] = (u16)(rl & 0xffffu); ln.readLength[1] = (u16)(rl >> 16);
            ln.R0 = rp; ln.R2 = rp + rs;
            ln.R = Str == 0 ? ln.R0 : ln.R2;
            // ---- the include/exclude recursion of stitchWindowAligns.cpp:8-353 for this sub-tree, warp-uniform, DFS cursor in registers.
            // Same visiting order and node accounting as dfsStep/dfsBacktrack (stitch.cu); the per-level phase array is replaced by two
            // bit masks: incl (seeds on the path) and open (included seeds whose exclude branch is still unexplored, i.e. not forced).
            DevTr* const tcur = ln.cur;
            Frame* const stack = ln.stack;
            __syncwarp();
            if (lane < 20) ((u32*)&tcur->h)[lane] = 0;   // trA = *trInit (ReadAlign_stitchPieces.cpp:282-286): every field zero
            __syncwarp();
            u32 level = 0, nInc = 0;
            int Score = 0;
            u32 tR2 = 0;
            u64 tG2 = 0, incl = 0, open = 0;
            u32 nodes = 0, leaves = 0;
            u32 curBlock = FLAT_NONE, firstBlock = FLAT_NONE, nCand = 0, inBlock = 0;
            int taskBest = 0;
            Cand c0; c0.mask = 0; c0.trOff = FLAT_NONE; c0.score = 0; c0.iFrag = 0; c0.pad = 0;
            #pragma unroll 1
            for (;;) {
                const u32 L = level;
                nodes++;
                if (L >= nA) {
                    if (tR2 != 0) {   // "iA>=nA && tR2==0: no aligns in the transcript" (:14)
                        leaves++;
                        const bool keep = evalLeaf<true>(ln, Score, tR2, tG2, Chr, Str, Str);
                        __syncwarp();
                        if (keep) {
                            const int sc = ln.leaf->h.maxScore;
                            Cand c; c.mask = incl; c.score = (short)sc; c.iFrag = ln.leaf->h.iFrag; c.pad = 0; c.trOff = FLAT_NONE;
                            if (fa.storeAll || sc + P.outFilterMultimapScoreRange >= taskBest) {   // likely to be recorded: keep the evaluated transcript
                                if (sc > taskBest) taskBest = sc;
                                const u32 nEx = ln.leaf->h.nExons;
                                const u32 words = (u32)(sizeof(TrHead) / 8) + nEx * (u32)(sizeof(Exon) / 8);
                                if (trCur + words > trEnd) {
                                    u64 off = 0;
                                    if (lane == 0) off = atomicAdd(&fa.bumps[3], (unsigned long long)(4 * FLAT_TR_CHUNK));
                                    off = __shfl_sync(0xffffffffu, off, 0);
                                    if (off + 4 * FLAT_TR_CHUNK <= fa.trWords && off + 4 * FLAT_TR_CHUNK < 0xFFFFFFFFULL) { trCur = off; trEnd = off + 4 * FLAT_TR_CHUNK; }
                                    else { trCur = 0; trEnd = 0; }
                                }
                                if (trCur + words <= trEnd) {
                                    u64* dst = fa.trStore + trCur;
                                    const u64* src = (const u64*)ln.leaf;   // head and exons are contiguous in DevTr
                                    #pragma unroll 1
                                    for (u32 q = lane; q < words; q += 32) dst[q] = src[q];
                                    c.trOff = (u32)trCur;
                                    trCur += words;
                                }
                            }
                            if (nCand == 0) {
                                c0 = c;
                                nCand = 1;
                            } else {
                                bool ok = true;
                                if (curBlock == FLAT_NONE || inBlock == FLAT_CAND_PER_BLOCK) {
                                    u64 nb = 0;
                                    if (lane == 0) nb = atomicAdd(&fa.bumps[2], 1ULL);
                                    nb = __shfl_sync(0xffffffffu, nb, 0);
                                    if (nb >= fa.maxBlocks) { if (lane == 0) fa.recs[k].over = 5; ok = false; }
                                    else {
                                        if (lane == 0) {
                                            fa.blocks[nb].next = FLAT_NONE; fa.blocks[nb].count = 0;
                                            if (curBlock != FLAT_NONE) fa.blocks[curBlock].next = (u32)nb;
                                        }
                                        if (curBlock == FLAT_NONE) firstBlock = (u32)nb;
                                        curBlock = (u32)nb;
                                        inBlock = 0;
                                    }
                                }
                                if (ok) {
                                    if (lane == 0) { FlatBlock& B = fa.blocks[curBlock]; B.c[inBlock] = c; B.count = inBlock + 1; }
                                    inBlock++;
                                    nCand++;
                                }
                            }
                        }
                    }
                    // unwind to the deepest included seed whose exclude branch is unexplored (undoing its include)
                    if (open == 0) break;
                    const u32 Bk = 63u - (u32)__clzll((long long)open);
                    const Frame& u = stack[--nInc];
                    if (u.h.nExons > 0) warpCopyWords(&tcur->ex[u.h.nExons - 1], &u.last, 6);
                    warpCopyWords(&tcur->h, &u.h, 20);
                    Score = u.Score; tR2 = u.tR2; tG2 = u.tG2;
                    open &= ~(1ULL << Bk); incl &= ~(1ULL << Bk);
                    level = Bk + 1;
                    continue;
                }
                const bool forced = L < depth;
                if (forced && ((bits >> (depth - 1 - L)) & 1u)) { level = L + 1; continue; }   // this level is fixed to "exclude"
                const u32 nEx0 = tcur->h.nExons;
                if (nEx0 > 0 && !forced) {
                    // The most frequent outcomes of an include attempt change nothing: the transcript is full (:13) or seed B ends inside
                    // the last included seed in read or genome space (:53-54, after the sjdb shortcut :18).  The test for seeds L, L+1, ..
                    // does not depend on the outcome for the earlier ones, so 32 seeds are tested at once.
                    const Exon& eA = tcur->ex[nEx0 - 1];
                    const bool full = nEx0 >= STAR_MAX_N_EXONS;
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
                                const bool sjdbDirect = q.sjA != SJA_NONE && eSj == q.sjA && (u64)q.rStart == (u64)tR2 + 1 && tG2 + 1 < q.gStart;
                                qf = !sjdbDirect && ((u64)q.rStart + q.Length - 1 <= tR2 || q.gStart + q.Length - 1 <= tG2);
                            }
                            real = !qf;
                        }
                        const u32 realMask = __ballot_sync(0xffffffffu, real);
                        const u32 span = nA - j < 32 ? nA - j : 32;
                        j += realMask ? (u32)__ffs(realMask) - 1 : span;
                        if (realMask || j >= nA) break;
                    }
                    if (j > L) { nodes += j - L - 1; level = j; continue; }
                }
                const Seed s = WA[L];
                if (nEx0 > 0 && forced) {   // a forced include that cannot succeed: the fixed prefix is not a valid path, the sub-tree is empty
                    const Exon& eA = tcur->ex[nEx0 - 1];
                    bool qf = nEx0 >= STAR_MAX_N_EXONS;
                    if (!qf && eA.iFrag == s.iFrag) {
                        const bool sjdbDirect = s.sjA != SJA_NONE && eA.sjA == s.sjA && (u64)s.rStart == (u64)tR2 + 1 && tG2 + 1 < s.gStart;
                        qf = !sjdbDirect && ((u64)s.rStart + s.Length - 1 <= tR2 || s.gStart + s.Length - 1 <= tG2);
                    }
                    if (qf) break;
                }
                Frame& u = stack[nInc];
                warpCopyWords(&u.h, &tcur->h, 20);
                if (nEx0 > 0) warpCopyWords(&u.last, &tcur->ex[nEx0 - 1], 6);
                if (lane == 0) { u.Score = Score; u.tR2 = tR2; u.tG2 = tG2; }   // (shared state has ONE writer: lane 0; readers come after a __syncwarp)
                int dScore;
                if (nEx0 > 0) {
                    dScore = stitchAlignToTranscript<true>(ln, tR2, tG2, s.rStart, s.gStart, s.Length, s.iFrag, s.sjA, tcur);
                } else {
                    if (lane == 0) {
                        tcur->ex[0].R = s.rStart; tcur->h.rStart = s.rStart;
                        tcur->ex[0].G = s.gStart; tcur->h.gStart = s.gStart;
                        tcur->ex[0].L = s.Length; tcur->ex[0].iFrag = s.iFrag; tcur->ex[0].sjA = s.sjA;
                        tcur->ex[0].canon = 0; tcur->ex[0].annot = 0; tcur->ex[0].sjStr = 0; tcur->ex[0].shL = 0; tcur->ex[0].shR = 0;
                        tcur->h.nExons = 1;
                        tcur->h.nMatch = s.Length;
                    }
                    dScore = s.Length;
                }
                __syncwarp();
                if (dScore > -1000000) {
                    if (lane == 0) {
                        if (s.Nrep == 1) tcur->h.nUnique++;
                        if (s.Anchor > 0) tcur->h.nAnchor++;
                    }
                    __syncwarp();
                    incl |= 1ULL << L;
                    if (!forced) open |= 1ULL << L;   // a forced include never explores its exclude branch
                    nInc++;
                    Score += dScore; tR2 = (u32)s.rStart + s.Length - 1; tG2 = s.gStart + s.Length - 1;
                } else {
                    if (forced) break;
                    if (u.h.nExons > 0) warpCopyWords(&tcur->ex[u.h.nExons - 1], &u.last, 6);   // the failed attempt may have touched the last exon / head
                    warpCopyWords(&tcur->h, &u.h, 20);
                }
                level = L + 1;
            }
            if (lane == 0) {
                FlatOut o;
                o.c0 = c0; o.count = nCand; o.first = firstBlock; o.nodes = nodes; o.leaves = leaves;
                fa.outs[t] = o;
            }
        }
    }
    PROF_ADD(17, clock64() - tStart);
}

// host-side launcher (the kernels are templates over the occupancy target; engine_api.cu is another translation unit)
#ifndef STAR_CUDA_HOST_SHIM   // (kernel launches need nvcc; the host emulation of the tests calls the kernels directly)
void launch_flat_dfs(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const FlatArgs& fa, u32* counter, const Caps& caps) {
    const u32 smem = 4 * FLAT_WARP_SMEM;
    if (ctasPerSM <= 4) flat_dfs_warp_kernel<4><<<nSM * 4, 128, smem, stream>>>(ix, P, fa, counter, caps);
    else if (ctasPerSM == 5) flat_dfs_warp_kernel<5><<<nSM * 5, 128, smem, stream>>>(ix, P, fa, counter, caps);
    else if (ctasPerSM <= 6) flat_dfs_warp_kernel<6><<<nSM * 6, 128, smem, stream>>>(ix, P, fa, counter, caps);
    else flat_dfs_warp_kernel<8><<<nSM * 8, 128, smem, stream>>>(ix, P, fa, counter, caps);
}
#endif

// ---- warp-uniform recording kernel: one WARP per read.  All 32 lanes execute the order-dependent recording with identical state
// (no divergence: one instruction stream per warp instead of 32); the scan over the read's tasks is cooperative: 32 FlatOut records
// per step, only tasks that produced candidates are visited.  Windows without candidates are skipped: windowBegin has no lasting
// effect for them and its exit conditions only depend on the number of transcripts recorded so far, which they do not change.
#define FLAT_REC_SMEM (2 * (u32)sizeof(DevTr) + 2 * (u32)sizeof(Frame) + 64)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) flat_record_warp_kernel(const __grid_constant__ DevIndex ix, const __grid_constant__ star_params_t P, ReadInfo* __restrict__ info,
                                                                    u32 nRecs, u32* __restrict__ counter, u8* __restrict__ arenas, Caps caps,
                                                                    star_read_result_t* __restrict__ results, star_align_t* __restrict__ staged, FlatArgs fa) {
    extern __shared__ u8 smem[];
    const u32 lane = threadIdx.x & 31;
    const u32 warpInBlock = threadIdx.x >> 5;
    const u32 gwarp = blockIdx.x * (blockDim.x >> 5) + warpInBlock;
    u8* ws = smem + (size_t)warpInBlock * FLAT_REC_SMEM;
    Lane ln;
    ln.leaf = (DevTr*)ws;
    ln.cur = ln.leaf + 1;
    ln.stack = (Frame*)(ln.cur + 1);
    ln.ph = (u8*)(ln.stack + 2);
    ln.ix = &ix; ln.P = &P; ln.R0 = nullptr; ln.R2 = nullptr; ln.R = nullptr; ln.caps = caps;
    {   // per-WARP arena: compacted window Chr/Str | transcript pool | pointer arrays
        u8* a = arenas + (u64)gwarp * caps.arenaBytes;
        ln.win = (Window*)a; a += (u64)caps.maxW * sizeof(Window);
        ln.pool = (DevTr*)a; a += (u64)caps.maxTr * sizeof(DevTr);
        ln.trPtr = (u16*)a; a += (u64)caps.maxTr * sizeof(u16);
        ln.winBase = (u16*)a; a += (u64)caps.maxW * sizeof(u16);
        ln.winN = (u16*)a;
        ln.wa = nullptr;
    }
    ln.memo = nullptr; ln.memoMask = 0; ln.memoBase = 0; ln.memoHit = 0; ln.memoMiss = 0; ln.lastSeed = -1; ln.coop = 1;
    long long tStart = clock64();
    long long nReplay = 0;
    #pragma unroll 1
    for (u32 q = lane; q < caps.maxTr; q += 32) ln.trPtr[q] = (u16)q;   // slot table: only ever permuted, initialised once
    __syncwarp();
    #pragma unroll 1
    for (;;) {
        u32 k = 0;
        if (lane == 0) k = atomicAdd(counter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= nRecs) break;
        const FlatRec rec = fa.recs[k];
        if (rec.done) continue;
        const u32 i = rec.read;
        ReadInfo ri = info[i];
        readBegin(ln, ri);
        ln.saEnum = rec.saEnum;
        if (rec.over) {
            ln.overflow = rec.over;
        } else {
            const u8* rp = fa.pool + rec.poolOff;
            const u32 rs = flatReadStride(rec.Lread);
            ln.R0 = rp; ln.R2 = rp + rs;
            const FlatWin* fw = (const FlatWin*)(rp + 2 * (u64)rs);
            const Seed* fs = (const Seed*)(fw + rec.nWin);
            u64 ndP = 0, lvP = 0;          // per-lane partial sums of the sub-tree work counters
            u32 curW = FLAT_NONE;
            bool openWin = false, stop = false;
            u16* wTr = nullptr; u16 nWinTr = 0;
            u32 Chr = 0, Str = 0, nA = 0;
            const Seed* WA = nullptr;
            #pragma unroll 1
            for (u32 t0 = 0; t0 < rec.nTasks; t0 += 32) {
                const u32 idx = t0 + lane;
                FlatOut o; o.count = 0; o.first = FLAT_NONE; o.nodes = 0; o.leaves = 0; o.c0.mask = 0; o.c0.trOff = FLAT_NONE; o.c0.score = 0; o.c0.iFrag = 0; o.c0.pad = 0;
                u32 wv = 0;
                if (idx < rec.nTasks) {
                    o = fa.outs[(u64)rec.taskBase + idx];
                    wv = fa.tasks[(u64)rec.taskBase + idx].w;
                    ndP += o.nodes; lvP += o.leaves;
                }
                u32 m = stop ? 0u : __ballot_sync(0xffffffffu, o.count > 0);
                #pragma unroll 1
                while (m && !stop) {
                    const u32 j = (u32)__ffs(m) - 1;
                    m &= m - 1;
                    const u32 w = __shfl_sync(0xffffffffu, wv, j);
                    const u32 cnt = __shfl_sync(0xffffffffu, o.count, j);
                    const u32 first = __shfl_sync(0xffffffffu, o.first, j);
                    Cand c0;
                    c0.mask = __shfl_sync(0xffffffffu, o.c0.mask, j);
                    c0.trOff = __shfl_sync(0xffffffffu, o.c0.trOff, j);
                    {
                        const u32 pk = __shfl_sync(0xffffffffu, (u32)(unsigned short)o.c0.score | ((u32)(u8)o.c0.iFrag << 16), j);
                        c0.score = (short)(pk & 0xffffu); c0.iFrag = (signed char)(pk >> 16); c0.pad = 0;
                    }
                    if (w != curW) {
                        if (openWin) { windowEnd(ln, Chr, Str, wTr, nWinTr, lane == 0); openWin = false; }
                        const int rc = windowBegin(ln, wTr, nWinTr, lane == 0);   // (shared state: lane 0 is the only writer)
                        __syncwarp();
                        if (rc == 2) { ln.overflow = 3; stop = true; break; }
                        if (rc == 1) { stop = true; break; }
                        const FlatWin W = fw[w];
                        Chr = W.Chr; Str = W.Str; nA = W.nWA;
                        WA = fs + W.seedOff;
                        ln.R = Str == 0 ? ln.R0 : ln.R2;
                        curW = w; openWin = true;
                    }
                    u32 b = first, inBlock = 0;
                    #pragma unroll 1
                    for (u32 q = 0; q < cnt; q++) {
                        Cand c;
                        if (q == 0) c = c0;
                        else {
                            if (inBlock == FLAT_CAND_PER_BLOCK) { b = fa.blocks[b].next; inBlock = 0; }
                            c = fa.blocks[b].c[inBlock++];
                        }
                        if (c.iFrag >= 0 && ln.maxScoreMate[c.iFrag] < c.score) ln.maxScoreMate[c.iFrag] = c.score;
                        const int wBest = ln.pool[wTr[0]].h.maxScore;
                        if (c.score + P.outFilterMultimapScoreRange >= wBest ||
                            (c.iFrag >= 0 && c.score + P.outFilterMultimapScoreRange >= ln.maxScoreMate[c.iFrag])) {
                            if (nWinTr > caps.maxTr - ln.trNtotal - 1) { ln.overflow = 3; stop = true; break; }
                            if (c.trOff != FLAT_NONE) {   // transcript stored by the warp that evaluated the leaf
                                const u32 nEx = ((const TrHead*)(fa.trStore + c.trOff))->nExons;
                                warpCopyWords(ln.leaf, fa.trStore + c.trOff, 20 + 6 * nEx);
                                recordLeaf<true>(ln, wTr, &nWinTr);
                            } else {
                                nReplay++;
                                __syncwarp();
                                u32 ok = 0;
                                if (lane == 0) {   // re-evaluation of the path (only when the transcript store was exhausted): sequential code, one lane
                                    int Score; u32 tR2; u64 tG2;
                                    ok = replayPath(ln, WA, nA, c.mask, Score, tR2, tG2) && evalLeaf(ln, Score, tR2, tG2, Chr, Str, Str);
                                }
                                ok = __shfl_sync(0xffffffffu, ok, 0);
                                __syncwarp();
                                if (ok) recordLeaf<true>(ln, wTr, &nWinTr);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            if (openWin) windowEnd(ln, Chr, Str, wTr, nWinTr, lane == 0);
            #pragma unroll 1
            for (int off = 16; off > 0; off >>= 1) { ndP += __shfl_xor_sync(0xffffffffu, ndP, off); lvP += __shfl_xor_sync(0xffffffffu, lvP, off); }
            ln.nodes = ndP; ln.leaves = lvP;
        }
        __syncwarp();
        if (lane == 0) selectExport(ln, ri, i, 0, 0, results, staged, info);   // sequential code on the read's final state: one lane
        __syncwarp();
    }
    PROF_ADD(18, clock64() - tStart);
    PROF_ADD(20, nReplay);
}

#ifndef STAR_CUDA_HOST_SHIM   // (kernel launches need nvcc; the host emulation of the tests calls the kernels directly)
void launch_flat_record(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, ReadInfo* info, u32 nRecs, u32* counter,
                        u8* arenas, const Caps& caps, star_read_result_t* results, star_align_t* staged, const FlatArgs& fa) {
    const u32 smem = 4 * FLAT_REC_SMEM;
    if (ctasPerSM <= 2) flat_record_warp_kernel<2><<<nSM * 2, 128, smem, stream>>>(ix, P, info, nRecs, counter, arenas, caps, results, staged, fa);
    else if (ctasPerSM == 3) flat_record_warp_kernel<3><<<nSM * 3, 128, smem, stream>>>(ix, P, info, nRecs, counter, arenas, caps, results, staged, fa);
    else flat_record_warp_kernel<4><<<nSM * 4, 128, smem, stream>>>(ix, P, info, nRecs, counter, arenas, caps, results, staged, fa);
}
#endif

#ifndef STAR_CUDA_HOST_SHIM   // (kernel launches need nvcc; the host emulation of the tests calls the kernels directly)
void launch_flat_setup(int ctasPerSM, int nSM, u32 smem, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info,
                       const Piece* pieces, u32 nHeavy, const u32* heavyList, const u64* heavyOff, const u8* heavyPool, u32* counter, u8* arenas, const Caps& caps,
                       star_read_result_t* results, star_align_t* staged, u32 smemStride, const FlatArgs& fa, u32 kBase) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(flat_setup_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);   // (sm_100: 227 KB per CTA; the overflow tier's window table of 2048 x 20 B per warp needs 203 KB)
        cudaFuncSetAttribute(flat_setup_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(flat_setup_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    if (ctasPerSM <= 2) flat_setup_kernel<2><<<nSM * 2, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, nHeavy, heavyList, heavyOff, heavyPool, counter, arenas, caps, results, staged, smemStride, fa, kBase);
    else if (ctasPerSM == 3) flat_setup_kernel<3><<<nSM * 3, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, nHeavy, heavyList, heavyOff, heavyPool, counter, arenas, caps, results, staged, smemStride, fa, kBase);
    else flat_setup_kernel<4><<<nSM * 4, 128, smem, stream>>>(ix, P, reads, stride, info, pieces, nHeavy, heavyList, heavyOff, heavyPool, counter, arenas, caps, results, staged, smemStride, fa, kBase);
}
#endif
