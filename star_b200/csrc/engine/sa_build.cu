// sa_build.cu — C-ABI of the suffix-array build (include/star_b200.h: star_gpu_sa_build).  Algorithm and kernels: sa_build_impl.cuh;
// the sorting / scanning / compaction primitives are cub's (library code, like the radix sort of the mapping path).
// No CPU fallback: without a CUDA device the call fails.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "dev.cuh"

namespace starb {
void setLastError(const std::string& m);     // engine_api.cu
void countLaunches(unsigned n);
static int g_saSM = 148;
static cudaError_t g_saErr = cudaSuccess;
static inline void saNote(cudaError_t e) { if (e != cudaSuccess && g_saErr == cudaSuccess) g_saErr = e; }
static inline void* saAlloc(size_t bytes) { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { saNote(cudaErrorMemoryAllocation); return nullptr; } return p; }
static inline unsigned saGrid(unsigned long long count) {
    const unsigned long long want = (count + 255) / 256, cap = (unsigned long long)g_saSM * 8;   // grid-stride: a multiple of the SM count once the job is large
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}
struct SaMax { __device__ __forceinline__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; } };
static void saSortPairs(const u64* kIn, u64* kOut, const u32* vIn, u32* vOut, u64 n, int endBit) {
    size_t tb = 0;
    saNote(cub::DeviceRadixSort::SortPairs(nullptr, tb, kIn, kOut, vIn, vOut, (long long)n, 0, endBit));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceRadixSort::SortPairs(tmp, tb, kIn, kOut, vIn, vOut, (long long)n, 0, endBit));
    countLaunches(4);
    cudaFree(tmp);
}
static void saMaxScan(u32* a, u64 n) {
    size_t tb = 0;
    saNote(cub::DeviceScan::InclusiveScan(nullptr, tb, a, a, SaMax(), (long long)n));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceScan::InclusiveScan(tmp, tb, a, a, SaMax(), (long long)n));
    countLaunches(2);
    cudaFree(tmp);
}
static void saSelect(const u32* in, const u8* flag, u32* out, u64 n, u64* nSel) {
    size_t tb = 0;
    unsigned long long* dN = (unsigned long long*)saAlloc(8);
    if (!dN) return;
    saNote(cub::DeviceSelect::Flagged(nullptr, tb, in, flag, out, dN, (long long)n));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceSelect::Flagged(tmp, tb, in, flag, out, dN, (long long)n));
    unsigned long long h = 0;
    saNote(cudaMemcpy(&h, dN, 8, cudaMemcpyDeviceToHost));
    *nSel = h;
    countLaunches(2);
    cudaFree(tmp); cudaFree(dN);
}
struct SaMax64 { __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; } };
static void saSortPairs64(const u64* kIn, u64* kOut, const u64* vIn, u64* vOut, u64 n, int endBit) {
    size_t tb = 0;
    saNote(cub::DeviceRadixSort::SortPairs(nullptr, tb, kIn, kOut, vIn, vOut, (long long)n, 0, endBit));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceRadixSort::SortPairs(tmp, tb, kIn, kOut, vIn, vOut, (long long)n, 0, endBit));
    countLaunches(4);
    cudaFree(tmp);
}
static void saMaxScan64(u64* a, u64 n) {
    size_t tb = 0;
    saNote(cub::DeviceScan::InclusiveScan(nullptr, tb, a, a, SaMax64(), (long long)n));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceScan::InclusiveScan(tmp, tb, a, a, SaMax64(), (long long)n));
    countLaunches(2);
    cudaFree(tmp);
}
template <class F> static void saSelectIf(F f, u64 lo, u64 hi, u64* out, u64* nSel) {   // values v in [lo, hi) with f(v), in order
    size_t tb = 0;
    *nSel = 0;
    if (hi <= lo) return;
    unsigned long long* dN = (unsigned long long*)saAlloc(8);
    if (!dN) return;
    thrust::counting_iterator<unsigned long long> it(lo);
    saNote(cub::DeviceSelect::If(nullptr, tb, it, out, dN, (long long)(hi - lo), f));
    void* tmp = saAlloc(tb);
    if (tmp) saNote(cub::DeviceSelect::If(tmp, tb, it, out, dN, (long long)(hi - lo), f));
    unsigned long long h = 0;
    saNote(cudaMemcpy(&h, dN, 8, cudaMemcpyDeviceToHost));
    *nSel = h;
    countLaunches(2);
    cudaFree(tmp); cudaFree(dN);
}
}  // namespace starb

#define SA_SORT_PAIRS64(kIn, kOut, vIn, vOut, n, endBit) starb::saSortPairs64(kIn, kOut, vIn, vOut, n, endBit)
#define SA_MAX_SCAN64(a, n) starb::saMaxScan64(a, n)
#define SA_SELECT_IF(f, lo, hi, out, nSel) starb::saSelectIf(f, lo, hi, out, nSel)
#define SA_ALLOC(bytes) starb::saAlloc(bytes)
#define SA_FREE(p) cudaFree(p)
#define SA_LAUNCH(count, kernel, ...) do { kernel<<<starb::saGrid(count), 256>>>(__VA_ARGS__); starb::saNote(cudaGetLastError()); starb::countLaunches(1); } while (0)
#define SA_LAUNCH_PACK(nRows, bits, kernel, ...)                                                                                   \
    do {   /* 4 warps per CTA, one tile of 32 x 64 rows per warp and step, tile staged in dynamic shared memory */              \
        const unsigned long long tiles_ = (((nRows) + 63) / 64 + 31) / 32, want_ = (tiles_ + 3) / 4, cap_ = (unsigned long long)starb::g_saSM * 8; \
        const size_t smem_ = (size_t)4 * 32 * (bits) * 8;                                                                          \
        starb::saNote(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));                      \
        kernel<<<(unsigned)(want_ < cap_ ? (want_ ? want_ : 1) : cap_), 128, smem_>>>(__VA_ARGS__);                                 \
        starb::saNote(cudaGetLastError());                                                                                         \
        starb::countLaunches(1);                                                                                                   \
    } while (0)
#define SA_COPY_TO(dst, src, bytes) starb::saNote(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice))
#define SA_COPY_FROM(dst, src, bytes) starb::saNote(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost))
#define SA_SORT_PAIRS(kIn, kOut, vIn, vOut, n, endBit) starb::saSortPairs(kIn, kOut, vIn, vOut, n, endBit)
#define SA_MAX_SCAN(a, n) starb::saMaxScan(a, n)
#define SA_SELECT(in, flag, out, n, nSel) starb::saSelect(in, flag, out, n, nSel)
#define SA_SYNC() starb::saNote(cudaDeviceSynchronize())
#define SA_ZERO(p, bytes) starb::saNote(cudaMemset(p, 0, bytes))
#include "sa_build_large.cuh"

using namespace starb;

extern "C" int star_gpu_sa_build(int device, const uint8_t* G, uint64_t nGenome, uint32_t GstrandBit, uint64_t nSA, uint8_t* SA, uint64_t nSAbyte) {
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0) {
        setLastError(std::string("star_b200: no CUDA device available (") + cudaGetErrorString(e) + "); index generation has no CPU fallback");
        return STAR_EXIT_RUNTIME;
    }
    if (device < 0 || device >= nDev) { setLastError("star_b200: bad device ordinal"); return STAR_EXIT_RUNTIME; }
    // texts beyond 32-bit ranks take the batched 64-bit path (sa_build_large.cuh); STAR_B200_SA_LARGE_CAP=<elements per sort> forces it (tests)
    u64 largeCap = 0;
    if (const char* e = getenv("STAR_B200_SA_LARGE_CAP")) largeCap = strtoull(e, nullptr, 10);
    if (2 * nGenome >= (1ULL << 32) - 64 && largeCap == 0) largeCap = 700000000ULL;
    if (2 * nGenome >= (1ULL << 33)) {
        setLastError("star_b200: the suffix-array build handles genomes up to 2^32 bases incl. padding");
        return STAR_EXIT_PARAMETER;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) g_saSM = prop.multiProcessorCount;
    g_saErr = cudaSuccess;
    u8* dG = (u8*)saAlloc(nGenome);
    const u64 outWords = (nSA + 63) / 64 * (GstrandBit + 1) + 2;
    u64* dOut = nullptr;
    int rc = 3;
    u64 rounds = 0;
    const bool timing = getenv("STAR_B200_SA_DEBUG") != nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (dG) {
        saNote(cudaMemcpy(dG, G, nGenome, cudaMemcpyHostToDevice));
        cudaEventRecord(e0);
        if (largeCap) {   // releases dG once the text exists, allocates the packed output after the ranks are gone (peak memory)
            rc = saBuildRunLarge(dG, nGenome, GstrandBit, nSA, &dOut, outWords, largeCap, &rounds);
            dG = nullptr;
        } else {
            dOut = (u64*)saAlloc(outWords * 8);
            if (dOut) {
                saNote(cudaMemset(dOut, 0, outWords * 8));
                rc = saBuildRun(dG, nGenome, GstrandBit, nSA, dOut, &rounds);
            }
        }
        cudaEventRecord(e1);
        if (rc == 0 && nSAbyte <= outWords * 8) saNote(cudaMemcpy(SA, dOut, nSAbyte, cudaMemcpyDeviceToHost));
        if (timing) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); fprintf(stderr, "star_b200 sa_build: n=%llu %s path, %llu rounds, %.1f ms on the device, rc %d\n", (unsigned long long)(2 * nGenome), largeCap ? "batched 64-bit" : "32-bit", (unsigned long long)rounds, ms, rc); }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dG); cudaFree(dOut);
    if (g_saErr != cudaSuccess) { setLastError(std::string("CUDA error in the suffix-array build: ") + cudaGetErrorString(g_saErr)); return g_saErr == cudaErrorMemoryAllocation ? STAR_EXIT_MEMORY_ALLOCATION : STAR_EXIT_RUNTIME; }
    if (rc == 3) { setLastError("star_b200: out of device memory for the suffix-array build"); return STAR_EXIT_MEMORY_ALLOCATION; }
    if (rc == 4) { setLastError("star_b200: suffix-array build: a 4-mer bin or a group of tied suffixes exceeds the sort capacity (STAR_B200_SA_LARGE_CAP)"); return STAR_EXIT_PARAMETER; }
    if (rc) { setLastError(rc == 1 ? "star_b200: suffix-array build: number of bases differs from nSA" : "star_b200: suffix-array build did not converge"); return STAR_EXIT_BUG; }
    return 0;
}
