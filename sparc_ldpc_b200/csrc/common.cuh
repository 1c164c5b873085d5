// common.cuh -- shared helpers for libsparc_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/sparc_b200.h"

namespace sb {

extern thread_local char g_err[512];
extern std::atomic<long> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define SB_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            snprintf(sb::g_err, sizeof(sb::g_err), "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                     cudaGetErrorString(e__));                                                  \
            return e__ == cudaErrorMemoryAllocation ? SB_ENOMEM : SB_ECUDA;                     \
        }                                                                                       \
    } while (0)

#define SB_LAUNCHED()                                   \
    do {                                                \
        sb::g_launches.fetch_add(1);                    \
        SB_CUDA(cudaGetLastError());                    \
    } while (0)

// Tuning knobs read from the environment exist only in experiment builds (make EXTRA=-DSB_EXPERIMENT,
// tools/ab_build*.sh): a release build ignores them, so a stray variable cannot change the grid or the
// shared-memory size of a kernel.
inline const char *knob(const char *name) {
#ifdef SB_EXPERIMENT
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

inline int ilog2(int v) {
    int r = 0;
    while ((1 << r) < v) r++;
    return r;
}

// Deterministic CTA-wide sum: xor-shuffle tree inside each warp, then thread 0 adds the warp
// partials in warp order.  `scratch` needs >= 33 doubles.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = scratch[0];
        for (int i = 1; i < nw; i++) s += scratch[i];
        scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

__device__ __forceinline__ double block_max(double v, double *scratch) {
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = scratch[0];
        for (int i = 1; i < nw; i++) s = fmax(s, scratch[i]);
        scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}

}  // namespace sb

struct sb_operator {
    int L, M, n, logM, w, H, Hp, NB;  // H = w/M blocks, Hp = padded to a multiple of 16, NB = Hp/16
    int pre, G8;                      // pre: inv entries are byte offsets; G8 = ceil(L/8)
    uint16_t *fwd;                    // [L][n]      lo | sign<<15                    (A beta gather, section lists)
    uint16_t *fwd8;                   // [G8][n][8]  the same entries, 8 sections interleaved (A beta gather, all sections)
    uint16_t *inv;                    // [L][M][Hp]  k (or k*8) / n (or n*8) = the zero word, visit order (A^T z fold)
    // FAST mode (fixed-point gathers): the same two maps with the order of every lane's terms chosen so that the
    // 32 lanes of a warp read distinct shared-memory banks (sched.h)
    int qok, qpre, qneg;  // qok: tables usable; qpre: invq holds byte offsets; qneg: word offset of the -z copy
    int PW, GQ;           // gather pool = PW (8 | 16) sections; GQ = floor(L / PW) chunks in fwdq
    uint16_t *invq;       // [L][EPT][Hp/8][TEAM][8]  word (or byte) offset into [ +z (n) | 32 zero words | -z (n) ]
    uint16_t *fwdq;       // [GQ][PW/8][n][8]         byte offset into the +-F area of a PW-section chunk (logM <= 9)
    // pair kernel (amp2.cu; M = 512, w/M <= 16, n <= 4608, L % 8 == 0): two codewords per CTA share every table word
    int p2ok;
    uint16_t *inv2;       // [L][16][2 signs][32][8]  byte offset into one codeword's z plane [n | 32 zero words]
    uint16_t *fwd2;       // [L/8][n][8]              slot*4096 + word*4 | sign << 15 inside an 8-section group buffer
    // SB_AMP_F64: the same two tables for 8-byte elements (k*8, slot*4096 + word*8), scheduled for half-warp pools of
    // 16 lanes x 16 eight-byte banks; built on first use from h_ordering (p2d_state: 0 = not yet, 1 = ready, -1 = none)
    uint16_t *inv2d, *fwd2d;
    uint32_t *h_ordering;
    int p2d_state;
};

struct sb_graph {
    int Nv, Nc, Nmsg, dcmax, dvmax;
    int *voff;     // [Nv+1]  prefix sums of vdeg
    int *vpos;     // [Nmsg]  internal message slot of each variable port
    int *cbase;    // [Nc]    internal slot of port 0 of each check
    int *cstride;  // [Nc]    slot stride between ports (= size of the check's degree class)
    int *cdeg;     // [Nc]
    int *ext2int;  // [Nmsg]  reference (check-major) message index -> internal slot
};
