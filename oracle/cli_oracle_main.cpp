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
    return star_cli_main_engine(argc, argv, &vt);
}
