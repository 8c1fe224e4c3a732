/*
 * star_oracle.h — entry points of the CPU restatement (oracle/star_oracle.cpp).
 * TEST INFRASTRUCTURE ONLY: see the header of star_oracle.cpp.
 */
#ifndef STAR_ORACLE_H
#define STAR_ORACLE_H
#include "../include/star_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* per-read intermediate state for unit-level parity tests (the goldens the reference never had) */
typedef struct star_oracle_dump {
    uint64_t* pcOff;   /* nReads+1 : first piece of read i */
    uint64_t* pc;      /* pieces PC[][8] after seeding (IncludeDefine.h:181-189): rStart,Length,Str,Dir,Nrep,SAstart,SAend,iFrag */
} star_oracle_dump_t;

int star_oracle_init(void** ctx, int device, const star_index_view_t* index, const star_params_t* params, uint32_t maxReads);
int star_oracle_map_chunk(void* ctx, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* stats);
int star_oracle_map_chunk_dump(void* ctx, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* stats,
                               star_oracle_dump_t* dump);
void star_oracle_dump_free(star_oracle_dump_t* d);
void star_oracle_destroy(void* ctx);
const char* star_oracle_last_error(void);
/* test-only: statistics of the k-ary search design check (env STAR_ORACLE_KARY_CHECK=1) */
void star_oracle_kary_stats(uint64_t* checked, uint64_t* mismatch);
void star_oracle_kary_cost(uint64_t* rounds, uint64_t* probes);
void star_oracle_window_hist(uint64_t* windows16, uint64_t* nodes16);
/* junction insertion: CPU restatement of star_gpu_sjdb_* (same arguments; include/star_b200.h) */
int star_oracle_sjdb_open(void** h, int device, const star_index_view_t* oldIndex);
int star_oracle_sjdb_search(void* h, const uint8_t* Gsj, uint64_t sjdbN, uint64_t sjdbLength, const uint8_t* skipSeq, uint64_t* indArray);
int star_oracle_sjdb_merge_sa(void* h, const uint64_t* indSorted, uint64_t nInd, uint64_t nGsj, uint64_t nGsjNew, uint64_t sjdbLength,
                              const uint32_t* oldSJind, uint8_t* SAnew, uint64_t nSAnewByte);
void star_oracle_sjdb_close(void* h);
/* 2nd stage of --outFilterType BySJout (same meaning as star_gpu_set_sj_novel) */
int star_oracle_set_sj_novel(void* ctx, const uint64_t* sjStart, const uint64_t* sjEnd, uint64_t n);
/* index generation: CPU restatement of star_gpu_sa_build */
int star_oracle_sa_build(int device, const uint8_t* G, uint64_t nGenome, uint32_t GstrandBit, uint64_t nSA, uint8_t* SA, uint64_t nSAbyte);
const star_engine_vtbl_t* star_oracle_engine(void);

#ifdef __cplusplus
}
#endif
#endif
