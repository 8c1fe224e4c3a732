// sjdb.cu — C-ABI of the device part of the on-the-fly junction insertion (include/star_b200.h: star_gpu_sjdb_*).
// Kernels: sjdb_kernels.cuh.  No CPU fallback: without a CUDA device star_gpu_sjdb_open fails.
#include <string>
#include <vector>

#include "sjdb_kernels.cuh"

namespace starb {
void setLastError(const std::string& m);     // engine_api.cu
void countLaunches(unsigned n);
}
using namespace starb;

struct star_sjdb {
    int device = 0;
    int nSM = 148;
    u8* dG = nullptr;       // 256 + nGenome + 256 bytes
    u64* dSA = nullptr;
    SjdbIndex ix;
    u64 sjGstart = 0, sjdbNold = 0;
};

#define SJ_CK(call)                                                                                                              \
    do {                                                                                                                         \
        cudaError_t e_ = (call);                                                                                                 \
        if (e_ != cudaSuccess) {                                                                                                 \
            setLastError(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
            return STAR_EXIT_RUNTIME;                                                                                            \
        }                                                                                                                        \
    } while (0)

extern "C" {

int star_gpu_sjdb_open(star_sjdb_t** out, int device, const star_index_view_t* v) {
    *out = nullptr;
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0) {
        setLastError(std::string("star_b200: no CUDA device available (") + cudaGetErrorString(e) + "); junction insertion has no CPU fallback");
        return STAR_EXIT_RUNTIME;
    }
    if (device < 0 || device >= nDev) { setLastError("star_b200: bad device ordinal"); return STAR_EXIT_RUNTIME; }
    SJ_CK(cudaSetDevice(device));
    star_sjdb* h = new star_sjdb;
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->nSM = prop.multiProcessorCount;
    const size_t gBytes = 256 + v->nGenome + 256, saWords = (v->nSAbyte + 7) / 8 + 2;
    if (cudaMalloc(&h->dG, gBytes) != cudaSuccess || cudaMalloc(&h->dSA, saWords * 8) != cudaSuccess) {
        setLastError("star_b200: out of device memory for the junction insertion");
        star_gpu_sjdb_close(h);
        return STAR_EXIT_MEMORY_ALLOCATION;
    }
    SJ_CK(cudaMemcpy(h->dG, v->G - 256, gBytes, cudaMemcpyHostToDevice));
    SJ_CK(cudaMemset(h->dSA, 0, saWords * 8));
    SJ_CK(cudaMemcpy(h->dSA, v->SA, v->nSAbyte, cudaMemcpyHostToDevice));
    h->ix.G = h->dG + 256; h->ix.SA = h->dSA; h->ix.nGenome = v->nGenome; h->ix.nSA = v->nSA;
    h->ix.GstrandBit = v->GstrandBit; h->ix.saBits = v->GstrandBit + 1;
    h->sjGstart = v->chrStart[v->nChrReal];
    h->sjdbNold = v->sjdbN;
    *out = h;
    return 0;
}

int star_gpu_sjdb_search(star_sjdb_t* h, const uint8_t* Gsj, uint64_t sjdbN, uint64_t sjdbLength, const uint8_t* skipSeq, uint64_t* indArray) {
    SJ_CK(cudaSetDevice(h->device));
    const u64 nSeq = 2 * sjdbN, nSuf = nSeq * sjdbLength;
    u8 *dGsj = nullptr, *dSkip = nullptr;
    u64* dInd = nullptr;
    SJ_CK(cudaMalloc(&dGsj, nSuf + 1 + 256));
    SJ_CK(cudaMalloc(&dSkip, nSeq + 1));
    SJ_CK(cudaMalloc(&dInd, nSuf * 16 + 16));
    SJ_CK(cudaMemset(dGsj, 5, nSuf + 1 + 256));
    SJ_CK(cudaMemcpy(dGsj, Gsj, nSuf + 1, cudaMemcpyHostToDevice));
    SJ_CK(cudaMemcpy(dSkip, skipSeq, nSeq, cudaMemcpyHostToDevice));
    const u64 want = (nSuf + 255) / 256;
    const unsigned grid = (unsigned)(want < (u64)h->nSM * 8 ? (want ? want : 1) : (u64)h->nSM * 8);   // a multiple of the SM count once the job is large
    sjdb_search_kernel<<<grid, 256>>>(h->ix, dGsj, nSeq, sjdbLength, dSkip, dInd);
    countLaunches(1);
    SJ_CK(cudaGetLastError());
    SJ_CK(cudaMemcpy(indArray, dInd, nSuf * 16, cudaMemcpyDeviceToHost));
    cudaFree(dGsj); cudaFree(dSkip); cudaFree(dInd);
    return 0;
}

int star_gpu_sjdb_merge_sa(star_sjdb_t* h, const uint64_t* indSorted, uint64_t nInd, uint64_t nGsj, uint64_t nGsjNew, uint64_t sjdbLength,
                           const uint32_t* oldSJind, uint8_t* SAnew, uint64_t nSAnewByte) {
    SJ_CK(cudaSetDevice(h->device));
    const u64 nSAnew = h->ix.nSA + nInd;
    std::vector<u64> row(nInd + 1), val(nInd + 1);
    sjdbInsertedRows(indSorted, nInd, h->ix.nSA, nGsj, h->sjGstart, h->ix.GstrandBit, row.data(), val.data());
    u64 *dRow = nullptr, *dVal = nullptr, *dOut = nullptr;
    u32* dOld = nullptr;
    const u64 outWords = (nSAnew + 63) / 64 * h->ix.saBits + 2;
    SJ_CK(cudaMalloc(&dRow, (nInd + 1) * 8));
    SJ_CK(cudaMalloc(&dVal, (nInd + 1) * 8));
    SJ_CK(cudaMalloc(&dOld, (h->sjdbNold + 1) * 4));
    SJ_CK(cudaMalloc(&dOut, outWords * 8));
    SJ_CK(cudaMemcpy(dRow, row.data(), (nInd + 1) * 8, cudaMemcpyHostToDevice));
    SJ_CK(cudaMemcpy(dVal, val.data(), (nInd + 1) * 8, cudaMemcpyHostToDevice));
    if (h->sjdbNold) SJ_CK(cudaMemcpy(dOld, oldSJind, h->sjdbNold * 4, cudaMemcpyHostToDevice));
    SjdbMerge m;
    m.insRow = dRow; m.insVal = dVal; m.nInd = nInd; m.nSAnew = nSAnew;
    m.nGenomeOld = h->ix.nGenome; m.nGenomeNew = h->sjGstart + nGsj; m.sjGstart = h->sjGstart; m.sjdbLength = sjdbLength; m.sjdbNold = h->sjdbNold;
    m.nGsjNew = nGsjNew; m.oldSJind = dOld;
    const u64 want = ((nSAnew + 63) / 64 + 255) / 256;
    const unsigned grid = (unsigned)(want < (u64)h->nSM * 8 ? (want ? want : 1) : (u64)h->nSM * 8);
    sjdb_merge_sa_kernel<<<grid, 256>>>(h->ix, m, dOut);
    countLaunches(1);
    SJ_CK(cudaGetLastError());
    if (nSAnewByte > outWords * 8) { setLastError("star_b200: bad size of the new suffix array"); return STAR_EXIT_BUG; }
    SJ_CK(cudaMemcpy(SAnew, dOut, nSAnewByte, cudaMemcpyDeviceToHost));
    cudaFree(dRow); cudaFree(dVal); cudaFree(dOld); cudaFree(dOut);
    return 0;
}

void star_gpu_sjdb_close(star_sjdb_t* h) {
    if (!h) return;
    cudaFree(h->dG);
    cudaFree(h->dSA);
    delete h;
}

}  // extern "C"
