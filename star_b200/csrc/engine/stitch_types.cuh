// stitch_types.cuh — plain-data types shared by the stitching kernels (stitch.cu, stitch_flat.cuh) and their launcher
// (engine_api.cu): kernel argument blocks and the records the three flat kernels exchange through HBM.
#pragma once
#include "dev.cuh"

namespace starb {

// ---- hand-over of DFS-heavy reads from stitch_kernel to the warp-per-read / flat kernels
struct HeavyArgs {
    u8* pool; u64 poolBytes; unsigned long long* bump;   // export pool
    u64* readOff;                                         // per read: offset of its record in the pool
    u32* list; u32* count;                                // heavy read list (read ids) and its length
    u32 estLimit;                                         // a read is heavy when sum_w 2^min(nWA_w,20) exceeds this (0 = heavy path off)
};

struct HeavyScratch {     // per warp, in HBM
    u32 maxTasks, maxBlocks, maxWin;
    u32 trWords;              // capacity (8-byte words) of the per-warp stored-transcript buffer
    u32 memoSlots;            // stitch memo entries per warp (power of two; 0 = off)
    u32 splitMin;             // windows with more seeds than this are cut into 2^(nWA-splitMin) (max 256) prefix sub-trees
    u64 bytesPerWarp;
};

// 16-byte leaf candidate: include mask of the path, offset (8-byte words) of the evaluated transcript in the transcript store
// (0xFFFFFFFF = not stored: the recording kernel replays the path), final score, mate of the transcript (-1: both)
struct Cand { u64 mask; u32 trOff; short score; signed char iFrag; u8 pad; };

// ---- flattened heavy path (stitch_flat.cuh)
struct FlatRec {          // one per heavy read of the chunk
    u64 poolOff;          // record in the flat pool
    u32 read;             // read index in the chunk
    u32 nWin;             // windows with seeds (only those are exported)
    u32 taskBase, nTasks; // tasks [taskBase, taskBase+nTasks) of the global task array
    u32 over;             // overflow reason (0 = ok); also set by flat_dfs_kernel when the candidate pool is exhausted
    u32 done;             // 1: the read was finished by the setup kernel (early exits)
    u32 saEnum;
    u32 Lread;
    u32 mmMax;            // outFilterMismatchNmaxTotal of the read
    u16 readLength[2];
};
static_assert(sizeof(FlatRec) == 48, "engine_api.cu sizes the record array with 48-byte entries");

struct FlatWin { u32 Chr; u16 nWA; u8 Str; u8 depth; u32 seedOff; u32 taskStart; };   // 16 B
struct FlatTask { u32 k; u16 w; u16 bits; };                                           // k = 0xFFFFFFFF: hole (never written)
struct FlatOut { Cand c0; u32 count; u32 first; u32 nodes; u32 leaves; };              // 32 B per task
#define FLAT_CAND_PER_BLOCK 7
struct FlatBlock { u32 next; u32 count; Cand c[FLAT_CAND_PER_BLOCK]; u64 pad; };       // 128 B
static_assert(sizeof(FlatOut) == 32 && sizeof(FlatBlock) == 128 && sizeof(FlatTask) == 8 && sizeof(FlatWin) == 16, "flat layouts");
#define FLAT_NONE 0xFFFFFFFFu
#define FLAT_TR_CHUNK 512   // 8-byte words a lane reserves at a time in the stored-transcript buffer

struct FlatArgs {
    FlatRec* recs;
    u8* pool; u64 poolBytes;
    unsigned long long* bumps;     // [0] pool bytes, [1] tasks, [2] candidate blocks, [3] stored-transcript words
    FlatTask* tasks; FlatOut* outs; u64 maxTasks;
    FlatBlock* blocks; u32 maxBlocks;
    u64* trStore; u64 trWords;
    u32 maxTasksPerRead, splitMin;
    u32 storeAll, slabByPos;       // storeAll: keep the evaluated transcript of EVERY surviving leaf (no replays in the recording kernel);
                                   // slabByPos: the piece slab of list entry k is slab k (overflow tier), not the slab of its read id
};

// launchers defined next to the (templated) kernels in stitch_flat.cuh
void launch_flat_setup(int ctasPerSM, int nSM, u32 smem, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const u8* reads, u32 stride, ReadInfo* info,
                       const Piece* pieces, u32 nHeavy, const u32* heavyList, const u64* heavyOff, const u8* heavyPool, u32* counter, u8* arenas, const Caps& caps,
                       star_read_result_t* results, star_align_t* staged, u32 smemStride, const FlatArgs& fa, u32 kBase);
void launch_flat_dfs(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, const FlatArgs& fa, u32* counter, const Caps& caps);
void launch_flat_record(int ctasPerSM, int nSM, cudaStream_t stream, const DevIndex& ix, const star_params_t& P, ReadInfo* info, u32 nRecs, u32* counter,
                        u8* arenas, const Caps& caps, star_read_result_t* results, star_align_t* staged, const FlatArgs& fa);

}  // namespace starb
