// cuda_host_shim.h — TEST INFRASTRUCTURE ONLY.  Lets the UNMODIFIED kernel sources of star_b200/csrc/engine (seed.cu, stitch.cu,
// stitch_flat.cuh) compile as host C++: one emulated CTA whose threads are host threads.  threadIdx/blockIdx/... are thread-local,
// the warp collectives (__ballot_sync, __shfl*_sync, __reduce_*_sync, __syncwarp, __all/__any_sync) meet at a per-warp barrier and
// exchange values through a per-warp slot array, __syncthreads is a CTA barrier, atomics are GCC atomics on plain memory, dynamic
// shared memory is one array per (single) CTA.  Used by oracle/engine_emul.cpp to run the whole kernel pipeline of a chunk on the
// CPU and compare it with the oracle: the device LOGIC (lane roles, collectives, memory layouts, arenas) is checked without a GPU.
#pragma once
#define STAR_CUDA_HOST_SHIM 1
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>   // vector types, qualifier macros (empty for a host compiler)

#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

namespace cuda_shim {
struct Dim { unsigned x, y, z; };
struct WarpShared {
    pthread_barrier_t bar; unsigned long long slot[32];
    pthread_barrier_t gbar[4][16];   // sub-warp groups of 2 / 4 / 8 / 16 lanes (collectives whose mask names one aligned group)
};
struct CtaShared { pthread_barrier_t bar; WarpShared warp[32]; unsigned nThreads; };
extern thread_local Dim tIdx, bIdx, bDim, gDim;
extern thread_local CtaShared* cta;
inline WarpShared& myWarp() { return cta->warp[tIdx.x >> 5]; }
inline unsigned laneId() { return tIdx.x & 31; }
inline void warpWait() { pthread_barrier_wait(&myWarp().bar); }
// barrier of the lanes named by `mask`: the whole warp, or one aligned group of 2 / 4 / 8 / 16 lanes (the group the caller is in)
inline void maskWait(unsigned mask) {
    if (mask == 0xffffffffu) { warpWait(); return; }
    const int G = __builtin_popcount(mask), lg = __builtin_ctz((unsigned)G) - 1;
    pthread_barrier_wait(&myWarp().gbar[lg][__builtin_ctz(mask) / G]);
}
template <class T> inline unsigned long long toSlot(T v) { unsigned long long s = 0; static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes"); memcpy(&s, &v, sizeof(T)); return s; }
template <class T> inline T fromSlot(unsigned long long s) { T v; memcpy(&v, &s, sizeof(T)); return v; }
template <class T> inline T exchange(T v, unsigned src, unsigned mask = 0xffffffffu) {   // every lane (of the mask) publishes v, reads lane src
    WarpShared& w = myWarp();
    w.slot[laneId()] = toSlot(v);
    maskWait(mask);
    T r = fromSlot<T>(w.slot[src & 31]);
    maskWait(mask);
    return r;
}
}  // namespace cuda_shim

#define threadIdx (cuda_shim::tIdx)
#define blockIdx (cuda_shim::bIdx)
#define blockDim (cuda_shim::bDim)
#define gridDim (cuda_shim::gDim)

template <class T> inline T __ldg(const T* p) { return *p; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cuda_shim::maskWait(mask); }
inline void __syncthreads() { pthread_barrier_wait(&cuda_shim::cta->bar); }
inline void __threadfence_block() { __sync_synchronize(); }
inline long long clock64() { return 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
inline unsigned __ballot_sync(unsigned mask, bool p) {
    cuda_shim::WarpShared& w = cuda_shim::myWarp();
    w.slot[cuda_shim::laneId()] = p ? 1 : 0;
    cuda_shim::maskWait(mask);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) if ((mask >> i) & 1) m |= (unsigned)(w.slot[i] & 1) << i;
    cuda_shim::maskWait(mask);
    return m;
}
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const unsigned l = cuda_shim::laneId();
    return cuda_shim::exchange(v, (l & ~(unsigned)(width - 1)) + ((unsigned)src & (unsigned)(width - 1)), mask);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d) { unsigned l = cuda_shim::laneId(); return cuda_shim::exchange(v, l + d < 32 ? l + d : l, mask); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d) { unsigned l = cuda_shim::laneId(); return cuda_shim::exchange(v, l >= d ? l - d : l, mask); }
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) { (void)width; return cuda_shim::exchange(v, cuda_shim::laneId() ^ (unsigned)m, mask); }
inline int __reduce_max_sync(unsigned, int v) {
    cuda_shim::WarpShared& w = cuda_shim::myWarp();
    w.slot[cuda_shim::laneId()] = (unsigned long long)(long long)v;
    cuda_shim::warpWait();
    int m = (int)(long long)w.slot[0];
    for (int i = 1; i < 32; i++) { int x = (int)(long long)w.slot[i]; if (x > m) m = x; }
    cuda_shim::warpWait();
    return m;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    cuda_shim::WarpShared& w = cuda_shim::myWarp();
    w.slot[cuda_shim::laneId()] = v;
    cuda_shim::warpWait();
    unsigned s = 0;
    for (int i = 0; i < 32; i++) s += (unsigned)w.slot[i];
    cuda_shim::warpWait();
    return s;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(volatile unsigned long long* p, unsigned long long v) { return __atomic_exchange_n((unsigned long long*)p, v, __ATOMIC_SEQ_CST); }
