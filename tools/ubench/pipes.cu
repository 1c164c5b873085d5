// Micro-benchmark (experiments only): which SM resources do the AMP kernel's phases share on B200?
//   mode 0: all 16 warps: conflict-free LDS.32 loop            (the fold / gather inner loop)
//   mode 1: all 16 warps: fp64 FMA loop                        (transforms / softmax)
//   mode 2: all warps LDS loop, then all warps fp64 loop       (phase-aligned, what barriers produce)
//   mode 3: warps 0-7 two LDS loops, warps 8-15 two fp64 loops (same work as mode 2, different pipes concurrently)
//   mode 4: LDS loop + 512-byte-per-warp LDG.128 table stream from L2 (16 KB per 256 LDS, the fold's ratio)
//   mode 5: LDS loop while one thread streams the same bytes into shared memory with cp.async.bulk (TMA)
//   mode 6: mode 5 + every warp also reads the staged bytes back with LDS.128
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NT 512
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int lds_loop(const int *zs, int iters, int lane, int salt) {
    int s = 0;
    uint32_t a = (lane + salt) & 8191;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) s += zs[(a + 32 * j * 7) & 8191];
        a = (a + 32 * 113 + (s & 0)) & 8191;
    }
    return s;
}

__device__ __forceinline__ double f64_loop(int iters, double seed) {
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = seed + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = fma(x[j], 1.0000001, 0.5);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += x[j];
    return s;
}

__global__ void __launch_bounds__(NT, 1) k(int mode, int iters, const uint4 *tab, size_t tab_u4, double *out) {
    extern __shared__ __align__(128) unsigned char raw[];
    int *zs = reinterpret_cast<int *>(raw);                    // 32 KB
    unsigned char *stage = raw + 32768;                        // 2 x 16 KB ring
    uint64_t *bar = reinterpret_cast<uint64_t *>(raw + 32768 + 32768);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8192; i += NT) zs[i] = i;
    if (tid == 0 && mode >= 5) {
        for (int b = 0; b < 2; b++) {
            uint32_t ba = smem_u32(bar + b);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double r = 0;
    const size_t base = ((size_t)blockIdx.x * 4099) % (tab_u4 - 1024 * 64);
    if (mode == 0) r = lds_loop(zs, iters, lane, warp);
    else if (mode == 1) r = f64_loop(iters * 2, tid);
    else if (mode == 2) { r = lds_loop(zs, iters, lane, warp); __syncthreads(); r += f64_loop(iters * 2, tid); }
    else if (mode == 3) { if (warp < 8) r = lds_loop(zs, 2 * iters, lane, warp); else r = f64_loop(iters * 4, tid); }
    else if (mode == 4) {
        int s = 0;
        const uint4 *t = tab + base + warp * 1024 + lane;
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int i = 0; i < iters; i += 8) {  // 8 x 16 LDS per 2 x 512-byte loads: 128 LDS per KB
            uint4 a = __ldg(t), b = __ldg(t + 32);
            t += 64;
            if (((i >> 3) & 15) == 15) t -= 1024;
            s += lds_loop(zs, 8, lane, warp + i);
            acc.x ^= a.x ^ b.x; acc.y ^= a.y ^ b.w;
        }
        r = s + (double)(acc.x ^ acc.y);
    } else {
        // one elected thread streams 16 KB chunks (what 16 warps consume per 16 x 16 LDS steps) into a 2-slot ring
        const int nchunk = iters / 16;  // per chunk: every warp does 16 lds_loop iterations = 256 LDS = 1 KB of table per warp
        int s = 0;
        uint32_t ph[2] = {0, 0};
        if (tid == 0) {
            for (int c = 0; c < 2 && c < nchunk; c++) {
                uint32_t ba = smem_u32(bar + c), dst = smem_u32(stage + c * 16384);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(tab + base + (size_t)c * 1024), "r"(16384), "r"(ba) : "memory");
            }
        }
        for (int c = 0; c < nchunk; c++) {
            const int b = c & 1;
            uint32_t ba = smem_u32(bar + b);
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(ba), "r"(ph[b]) : "memory");
            }
            ph[b] ^= 1;
            if (mode == 6) {
                const uint4 *st = reinterpret_cast<const uint4 *>(stage + b * 16384) + warp * 64 + lane;
                uint4 a = st[0], bb = st[32];
                s += a.x ^ bb.y;
            }
            s += lds_loop(zs, 16, lane, warp + c);
            __syncthreads();
            if (tid == 0 && c + 2 < nchunk) {
                uint32_t dst = smem_u32(stage + b * 16384);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(tab + base + (size_t)((c + 2) & 63) * 1024), "r"(16384), "r"(ba) : "memory");
            }
        }
        r = s;
    }
    if (r == 123.456) out[blockIdx.x * NT + tid] = r;
}

int main() {
    const size_t tab_bytes = 64u << 20;
    uint4 *tab; double *out;
    cudaMalloc(&tab, tab_bytes); cudaMemset(tab, 1, tab_bytes);
    cudaMalloc(&out, 148 * NT * 8);
    const int smem = 32768 + 32768 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    const char *names[] = {"LDS only", "fp64 only", "LDS then fp64 (aligned phases)", "LDS || fp64 (warp halves)",
                           "LDS + LDG.128 table stream", "LDS + TMA bulk stream (not read)", "LDS + TMA bulk stream + LDS.128 read"};
    for (int mode = 0; mode < 7; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k<<<148, NT, smem>>>(mode, iters, tab, tab_bytes / 16, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        // cycles per 16-LDS step per SM (16 warps each doing one step = 256 warp-LDS)
        printf("mode %d  %-40s %8.3f ms  (%s)  %.1f cycles per (16 warps x 16 LDS) step @1.965 GHz\n", mode, names[mode], best,
               cudaGetErrorString(e), best * 1e-3 * 1.965e9 / iters);
    }
    return 0;
}
