// warp_prims.cuh — the warp collectives used by warp-uniform device code, behind one small interface.
//
// On the GPU (`DevWarp`) every member is the CUDA intrinsic.  With STAR_WARP_HOST_EMUL defined (tests only: oracle/warp_emul.cpp)
// the same source compiles for the host and `HostWarp` runs the 32 lanes as 32 threads that meet at a barrier in every collective,
// so that the exact device logic (lane roles, ballots, shuffles, reductions) can be checked against the oracle without a GPU.
// Code written against this interface must call the collectives from all 32 lanes (warp-uniform control flow around them).
#pragma once
#include <stdint.h>

#ifdef STAR_WARP_HOST_EMUL
#include <pthread.h>
#include <string.h>
#define SB_DEV inline
#define SB_LDG(p) (*(p))
#define SB_FFS(x) __builtin_ffs((int)(x))            /* 1-based index of the lowest set bit, 0 if none */
#define SB_CLZ(x) __builtin_clz((unsigned)(x))        /* x != 0 */
#define SB_CTZ(x) __builtin_ctz((unsigned)(x))        /* x != 0 */
#define SB_CTZ64(x) __builtin_ctzll((unsigned long long)(x))
#else
#define SB_DEV __device__ __forceinline__
#define SB_LDG(p) __ldg(p)
#define SB_FFS(x) __ffs((int)(x))
#define SB_CLZ(x) __clz((int)(x))
#define SB_CTZ(x) (__ffs((int)(x)) - 1)
#define SB_CTZ64(x) (__ffsll((long long)(x)) - 1)
#endif

namespace starb {

#ifndef STAR_WARP_HOST_EMUL
struct DevWarp {
    unsigned lane;
    SB_DEV DevWarp() : lane(threadIdx.x & 31) {}
    SB_DEV unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    SB_DEV unsigned shfl(unsigned v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    SB_DEV unsigned long long shfl64(unsigned long long v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    SB_DEV int reduceMax(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    SB_DEV unsigned reduceAdd(unsigned v) const { return __reduce_add_sync(0xffffffffu, v); }
    SB_DEV void sync() const { __syncwarp(); }
};
#else
// 32 host threads share one HostWarpShared; every collective is: publish my value, barrier, read, barrier.
struct HostWarpShared {
    pthread_barrier_t bar;
    unsigned long long slot[32];
    HostWarpShared() { pthread_barrier_init(&bar, nullptr, 32); memset(slot, 0, sizeof(slot)); }
    ~HostWarpShared() { pthread_barrier_destroy(&bar); }
};
struct HostWarp {
    unsigned lane;
    HostWarpShared* sh;
    HostWarp(unsigned l, HostWarpShared* s) : lane(l), sh(s) {}
    void wait() const { pthread_barrier_wait(&sh->bar); }
    unsigned ballot(bool p) const {
        sh->slot[lane] = p ? 1 : 0; wait();
        unsigned m = 0;
        for (int i = 0; i < 32; i++) m |= (unsigned)(sh->slot[i] & 1) << i;
        wait();
        return m;
    }
    unsigned long long shfl64(unsigned long long v, int src) const {
        sh->slot[lane] = v; wait();
        unsigned long long r = sh->slot[src & 31];
        wait();
        return r;
    }
    unsigned shfl(unsigned v, int src) const { return (unsigned)shfl64(v, src); }
    int reduceMax(int v) const {
        sh->slot[lane] = (unsigned long long)(long long)v; wait();
        int m = (int)(long long)sh->slot[0];
        for (int i = 1; i < 32; i++) { int x = (int)(long long)sh->slot[i]; if (x > m) m = x; }
        wait();
        return m;
    }
    unsigned reduceAdd(unsigned v) const {
        sh->slot[lane] = v; wait();
        unsigned s = 0;
        for (int i = 0; i < 32; i++) s += (unsigned)sh->slot[i];
        wait();
        return s;
    }
    void sync() const { wait(); }
};
#endif

}  // namespace starb
