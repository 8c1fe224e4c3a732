// seed_types.cuh — records exchanged by the kernels of the keyed seed stage (seed_keyed.cuh) and sized by the context (engine_api.cu).
#pragma once
#include "dev.cuh"

namespace starb {

struct ChainItem {              // one chain of searches: piece (ps, pl) of `read`, chainId = (piece << 8) | (direction << 7) | start
    u32 read;
    u16 ps, pl;
    u16 chainId;
    u8 iFrag, nStart;
};
struct SeedRec {                // result of one maxMappableLength2strands call, replayed through storeAligns
    u64 SAstart;
    u32 Nrep;
    u16 Shift, L;
    u16 chainId;
    u8 k;                       // position in the chain; 255 = the fixed-length search of --seedSearchLmax
    u8 nSai;                    // SAindex words read by this search
    u8 iFrag, pad_[3];
};
static_assert(sizeof(ChainItem) == 12 && sizeof(SeedRec) == 24, "seed stage record layout");

struct KeyedArgs {
    const u32* saKeys;          // one key per SA row (16-byte aligned, 16 bytes of slack behind)
    ChainItem* items;           // chains of the chunk (read == 0xffffffff: unused slot)
    u32* itemKey;               // sort key of every item (SAindex L-mer of the chain's first search)
    u32* itemIdx;               // 0, 1, 2, ... (the values of the key sort)
    u32* itemCount;             // [0] number of items
    u32 maxItems;
    SeedRec* recs;              // maxRec records per read
    u32* recCount;              // per read
    u32 maxRec;
    u32 scanMax;                // windows up to this many rows are scanned tile by tile; larger ones are bisected on the keys
};

}  // namespace starb
