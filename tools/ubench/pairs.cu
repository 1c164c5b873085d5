// Micro-benchmark (experiments only): which instruction classes run concurrently on a B200 SM?
// 16 warps per SM; warps 0-7 run work kind A, warps 8-15 run work kind B, each sized to take about the same time alone.
// kinds: 0 = LDS.32 conflict-free, 1 = fp64 FMA, 2 = int32 add/xor, 3 = fp32 FMA, 4 = LDS.128 (few ALU ops per wavefront)
// Prints time(A alone on 8 warps), time(B alone on 8 warps), time(A || B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NT 512

__device__ __forceinline__ double work(int kind, int iters, const int *zs, int lane, int warp) {
    if (kind == 0) {
        int s = 0;
        const int *p = zs + lane;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 16; j++) s += p[j * 224];
            p = zs + ((lane + (i & 7) * 32 + (s & 0)) & 1023);
        }
        return s;
    } else if (kind == 1) {
        double x[16];
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = lane + j;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 16; j++) x[j] = fma(x[j], 1.0000001, 0.5);
        }
        double s = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) s += x[j];
        return s;
    } else if (kind == 2) {
        int x[16];
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = lane * 77 + j;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 16; j++) x[j] = (x[j] + (x[(j + 1) & 15] ^ i));
        }
        int s = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) s ^= x[j];
        return s;
    } else if (kind == 3) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = lane + j;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 16; j++) x[j] = fmaf(x[j], 1.0000001f, 0.5f);
        }
        float s = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) s += x[j];
        return s;
    } else {
        int s = 0;
        const int4 *p = reinterpret_cast<const int4 *>(zs) + lane;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) { const int4 v = p[j * 32]; s ^= v.x ^ v.y ^ v.z ^ v.w; }
            p = reinterpret_cast<const int4 *>(zs) + ((lane + (i & 7) * 32 + (s & 0)) & 255);
        }
        return s;
    }
}

__global__ void __launch_bounds__(NT, 1) k(int kindA, int itA, int kindB, int itB, double *out) {
    __shared__ int zs[8192];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8192; i += NT) zs[i] = i;
    __syncthreads();
    double r = 0;
    if (warp < 8) { if (itA > 0) r = work(kindA, itA, zs, lane, warp); }
    else { if (itB > 0) r = work(kindB, itB, zs, lane, warp); }
    if (r == 123.456) out[blockIdx.x * NT + tid] = r;
}

static float run(int ka, int ia, int kb, int ib, double *out) {
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148, NT>>>(ka, ia, kb, ib, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    double *out; cudaMalloc(&out, 148 * NT * 8);
    const char *nm[] = {"LDS.32", "fp64 FMA", "int32 ALU", "fp32 FMA", "LDS.128"};
    const int base[] = {4096, 4096, 8192, 16384, 4096};
    float alone[5];
    for (int a = 0; a < 5; a++) {
        alone[a] = run(a, base[a], 0, 0, out);
        printf("%-10s alone on 8 warps: %.3f ms  (%.2f cycles per warp-instruction-group of 16)\n", nm[a], alone[a],
               alone[a] * 1e-3 * 1.965e9 / base[a]);
    }
    for (int a = 0; a < 5; a++)
        for (int b = a; b < 5; b++) {
            // scale B's iterations so that both halves take the same time alone
            const int ib = (int)(base[b] * alone[a] / alone[b]);
            const float tb = run(b, ib, 0, 0, out);
            const float both = run(a, base[a], b, ib, out);
            printf("%-10s || %-10s: alone %.3f / %.3f ms, together %.3f ms  -> overlap factor %.2f (1 = free, 2 = serialised)\n",
                   nm[a], nm[b], alone[a], tb, both, both / (0.5f * (alone[a] + tb)));
        }
    return 0;
}
