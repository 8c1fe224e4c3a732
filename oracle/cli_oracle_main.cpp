// cli_oracle_main.cpp — TEST-ONLY binary: the repository's host code (parameter parsing, index loading, FASTQ
// chunking, SAM/SJ/Log writers: star_b200/csrc/host/) driven by the CPU oracle instead of the CUDA engine.
// Used by tests/ to (a) pin the oracle + host code against the unmodified reference (oracle/_ref/STAR) and
// (b) produce expected outputs on machines without a GPU.  Never shipped, never linked into libstar_b200.so.
//
// With STAR_CLI_SJDB_EMUL=<path of libengine_emul.so> the device steps of the junction insertion and of the index generation are run
// by the EMULATED CUDA kernels (sjdb_kernels.cuh, sa_build_impl.cuh through cuda_host_shim.h) instead of the oracle's sequential
// restatements, so that a whole 2-pass run / genomeGenerate checks the kernel logic against the reference's outputs.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "star_oracle.h"

namespace {
typedef int (*emul_search_t)(const star_index_view_t*, const uint8_t*, uint64_t, uint64_t, const uint8_t*, uint64_t*);
typedef int (*emul_merge_t)(const star_index_view_t*, const uint64_t*, uint64_t, uint64_t, uint64_t, uint64_t, const uint32_t*, uint8_t*, uint64_t);
emul_search_t g_search;
emul_merge_t g_merge;
int emOpen(void** h, int, const star_index_view_t* v) { *h = (void*)v; return 0; }
int emSearch(void* h, const uint8_t* Gsj, uint64_t sjdbN, uint64_t L, const uint8_t* skip, uint64_t* ind) { return g_search((const star_index_view_t*)h, Gsj, sjdbN, L, skip, ind); }
int emMerge(void* h, const uint64_t* ind, uint64_t nInd, uint64_t nGsj, uint64_t nGsjNew, uint64_t L, const uint32_t* old, uint8_t* SAnew, uint64_t nByte) {
    return g_merge((const star_index_view_t*)h, ind, nInd, nGsj, nGsjNew, L, old, SAnew, nByte);
}
void emClose(void*) {}

// STAR_CLI_PINNED_EMUL=1: the optional vtable members of the CUDA engine (page-locked buffers, second fetch after a too small
// out->alignsCapacity) are emulated around the oracle engine so that the driver's use of them is tested without a GPU.
const star_engine_vtbl_t* g_base;
std::vector<star_align_t> g_lastAligns;
std::vector<star_read_result_t> g_lastReads;
long g_hostAllocs = 0, g_capacityMisses = 0;
void* pinAlloc(size_t n) { g_hostAllocs++; return malloc(n ? n : 1); }
void pinFree(void* p) { free(p); }
int pinDownload(void*, star_align_batch_t* out) {
    if (g_lastAligns.size() > out->alignsCapacity) { out->nAligns = g_lastAligns.size(); g_capacityMisses++; return STAR_EXIT_RUNTIME; }
    if (!g_lastReads.empty()) memcpy(out->reads, g_lastReads.data(), g_lastReads.size() * sizeof(star_read_result_t));
    if (!g_lastAligns.empty()) memcpy(out->aligns, g_lastAligns.data(), g_lastAligns.size() * sizeof(star_align_t));
    out->nAligns = g_lastAligns.size();
    return 0;
}
int pinMap(void* ctx, const star_read_batch_t* in, star_align_batch_t* out, star_chunk_stats_t* st) {
    g_lastReads.assign(in->nReads, star_read_result_t());
    g_lastAligns.resize((size_t)in->nReads * 64 + 64);
    star_align_batch_t tmp;
    tmp.reads = g_lastReads.data(); tmp.aligns = g_lastAligns.data(); tmp.alignsCapacity = g_lastAligns.size(); tmp.nAligns = 0;
    const int rc = g_base->map_chunk(ctx, in, &tmp, st);
    if (rc) return rc;
    g_lastAligns.resize(tmp.nAligns);
    return pinDownload(ctx, out);
}
}  // namespace

int main(int argc, char** argv) {
    star_engine_vtbl_t vt = *star_oracle_engine();
    if (const char* lib = getenv("STAR_CLI_SJDB_EMUL")) {
        void* so = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
        if (!so) { fprintf(stderr, "cannot load %s: %s\n", lib, dlerror()); return 1; }
        g_search = (emul_search_t)dlsym(so, "engine_emul_sjdb_search");
        g_merge = (emul_merge_t)dlsym(so, "engine_emul_sjdb_merge_sa");
        if (!g_search || !g_merge) { fprintf(stderr, "%s lacks the sjdb entry points\n", lib); return 1; }
        vt.sjdb_open = emOpen; vt.sjdb_search = emSearch; vt.sjdb_merge_sa = emMerge; vt.sjdb_close = emClose;
        typedef int (*emul_sa_t)(int, const uint8_t*, uint64_t, uint32_t, uint64_t, uint8_t*, uint64_t);
        emul_sa_t sa = (emul_sa_t)dlsym(so, "engine_emul_sa_build");
        if (!sa) { fprintf(stderr, "%s lacks engine_emul_sa_build\n", lib); return 1; }
        vt.sa_build = sa;
    }
    if (getenv("STAR_CLI_PINNED_EMUL")) {
        static star_engine_vtbl_t base = vt;
        g_base = &base;
        vt.map_chunk = pinMap; vt.host_alloc = pinAlloc; vt.host_free = pinFree; vt.download_results = pinDownload;
    }
    const int rc = star_cli_main_engine(argc, argv, &vt);
    if (getenv("STAR_CLI_PINNED_EMUL")) fprintf(stderr, "pinned emulation: %ld host allocations, %ld capacity misses\n", g_hostAllocs, g_capacityMisses);
    return rc;
}
