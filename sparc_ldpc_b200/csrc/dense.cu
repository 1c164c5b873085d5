// dense.cu -- Gaussian (dense) design-matrix mode of the AMP decoder: A beta and A^T z as tcgen05 / TMA GEMMs.
//
// The reference's amp() takes the design operator as two closures (ldpc/sparc_ldpc.py:189, :213, :220); with a
// dense i.i.d. Gaussian A (BASELINE configs[0]) they are s = beta + A^T z and A beta for every codeword of the
// batch, i.e. two GEMMs per AMP iteration:  [LM x n] x [n x B]  and  [n x LM] x [LM x B].
//
// Precision scheme (FP32 emulation on the bf16 tensor pipe, "bf16x3"): every operand value x (fp64) is split into
// three bf16 planes x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)  (|x - x0 - x1 - x2| <= 2^-27 |x|).
// A product a*b is evaluated as a0 b0 + (a0 b1 + a1 b0 + a0 b2 + a1 b1 + a2 b0): six tcgen05.mma passes whose
// bf16 x bf16 products are exact in the fp32 accumulators; the three dropped terms are <= 2^-26 |a b|.  The
// leading term and the five correction terms accumulate in two separate TMEM accumulators, so the corrections
// are not rounded at the magnitude of the leading sum, and the K loop is cut into chunks: after every chunk the
// epilogue warps drain both accumulators and add them in fp64 (hi + lo) to the output, which bounds the length
// of any fp32 accumulation to one chunk (default 1024 values of k).
//
// Kernel: one CTA per (128-row tile of the A-side operand, BN codewords, K slice).  Warp 0 = TMA producer (3 + 3
// bf16 planes per stage, 64-byte rows, SWIZZLE_64B), warp 1 = TMEM allocation + single-thread MMA issue
// (12 tcgen05.mma per 32-k stage), warps 2-5 = epilogue (tcgen05.ld 32x32b, fp64 accumulate into the partial
// output of the CTA's K slice).  Operands are K-major on both sides, so A is stored twice (A and A^T planes).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace sb {

constexpr int GM = 128;      // UMMA M: rows of the A-side operand per CTA
constexpr int GK = 32;       // k values per pipeline stage
constexpr int GKB = GK * 2;  // bytes per smem row (bf16) = the swizzle span

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// the same box, delivered to the same CTA-relative shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the barrier at the same offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major [rows][32 bf16] tile written by TMA with SWIZZLE_64B:
// start address >> 4 in bits [0,14), leading byte offset field = 1 (unused by swizzled K-major layouts) in bits
// [16,30), stride byte offset (8 rows * 64 B = 512 B) >> 4 in bits [32,46),
// descriptor version 1 in bits [46,48), layout type SWIZZLE_64B = 4 in bits [61,64)
// (cute/arch/mma_sm100_desc.hpp, union SmemDescriptor)
__device__ __forceinline__ uint64_t kmajor_sw64_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

// instruction descriptor: D = fp32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
}

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = 3 * GM * GKB, B_BYTES = 3 * BN * GKB, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 8 * (2 * STAGES + 2) + 16;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + slack for the 1024-byte alignment
};

// out[slice][nn][mm] (=|+=) sum_k Aop[mm][k] * Bop[nn][k]   over the k-blocks of the CTA's slice
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   double *__restrict__ out, int M, int N, int ldo, long slice_stride, int nkb, int kb_per_slice,
                   int kb_per_chunk, int CL) {
    // CL = thread-block cluster size along the row tiles (1 | 2 | 4): the CL CTAs of a cluster work on CL
    // consecutive 128-row tiles of the A-side operand against the SAME codeword tile and K slice, so each CTA
    // fetches only BN / CL rows of the codeword-side operand and TMA-multicasts them to all CL CTAs (the kernel
    // is bound by L2 -> SM traffic: 72 KB per 32-k stage alone, 48 KB with CL = 2, 36 KB with CL = 4).
    using S = GemmSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * S::STAGE_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tfull = bars + 8u * (2 * STAGES), tempty = tfull + 8u, tslot = tempty + 8u;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = reinterpret_cast<volatile uint32_t *>(gen_base + STAGES * S::STAGE_BYTES + 8 * (2 * STAGES + 2));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * BN, slice = blockIdx.z;
    const int kb0 = slice * kb_per_slice, kb1 = min(nkb, kb0 + kb_per_slice);

    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), CL);  // a stage is refilled by every CTA of the cluster: all of them must release it
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: [0, BN) leading-term accumulator, [BN, 2 BN) correction-term accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; kb++) {
                mbar_wait(empty(s), ph ^ 1u);
                mbar_expect_tx(full(s), S::STAGE_BYTES);
                const uint32_t dst = base + s * S::STAGE_BYTES;
#pragma unroll
                for (int p = 0; p < 3; p++) tma_load_3d(dst + p * GM * GKB, &tmA, full(s), kb * GK, m0, p);
                if (CL == 1) {
#pragma unroll
                    for (int p = 0; p < 3; p++) tma_load_3d(dst + S::A_BYTES + p * BN * GKB, &tmB, full(s), kb * GK, n0, p);
                } else {  // this CTA's share of the codeword rows, to every CTA of the cluster
                    const int rows = BN / CL;
#pragma unroll
                    for (int p = 0; p < 3; p++)
                        tma_load_3d_mc(dst + S::A_BYTES + p * BN * GKB + crank * rows * GKB, &tmB, full(s), kb * GK,
                                       n0 + crank * rows, p, cmask);
                }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issue (one thread)
            constexpr uint32_t idesc = umma_idesc(BN);
            const uint32_t d_hi = tmem, d_lo = tmem + BN;
            int s = 0, chunk = 0;
            uint32_t ph = 0, tph = 0;
            for (int kb = kb0; kb < kb1;) {
                if (chunk > 0) {  // the epilogue must have drained the previous chunk
                    mbar_wait(tempty, tph);
                    tph ^= 1u;
                    tc_fence_after();
                }
                const int kend = min(kb1, kb + kb_per_chunk);
                uint32_t acc = 0;
                for (; kb < kend; kb++) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t a_base = base + s * S::STAGE_BYTES, b_base = a_base + S::A_BYTES;
#pragma unroll
                    for (int j = 0; j < GK / 16; j++) {
                        uint64_t ad[3], bd[3];
#pragma unroll
                        for (int p = 0; p < 3; p++) {
                            ad[p] = kmajor_sw64_desc(a_base + p * GM * GKB + j * 32);
                            bd[p] = kmajor_sw64_desc(b_base + p * BN * GKB + j * 32);
                        }
                        umma_bf16(d_hi, ad[0], bd[0], idesc, acc);
                        umma_bf16(d_lo, ad[0], bd[1], idesc, acc);
                        umma_bf16(d_lo, ad[1], bd[0], idesc, 1u);
                        umma_bf16(d_lo, ad[0], bd[2], idesc, 1u);
                        umma_bf16(d_lo, ad[1], bd[1], idesc, 1u);
                        umma_bf16(d_lo, ad[2], bd[0], idesc, 1u);
                        acc = 1u;
                    }
                    if (CL == 1) umma_commit(empty(s));  // frees the stage once these MMAs have read it
                    else umma_commit_mc(empty(s), cmask);
                    if (++s == STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull);  // chunk complete -> epilogue
                chunk++;
            }
        }
    } else {  // ---- epilogue warps 2..5: TMEM lanes [32 g, 32 g + 32), g = warp % 4
        const int g = warp & 3, row = g * 32 + lane, m = m0 + row;
        const int nchunks = (kb1 - kb0 + kb_per_chunk - 1) / kb_per_chunk;
        double *obase = out + (size_t)slice * slice_stride;
        uint32_t fph = 0;
        for (int c = 0; c < nchunks; c++) {
            mbar_wait(tfull, fph);
            fph ^= 1u;
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; cb++) {
                uint32_t hi[32], lo[32];
                const uint32_t taddr = tmem + ((uint32_t)(g * 32) << 16) + cb * 32;
                tmem_ld32(taddr, hi);
                tmem_ld32(taddr + BN, lo);
                tmem_ld_wait();
                if (m < M) {
                    // the CTA owns this slice of the output: the first chunk stores, later chunks add with a
                    // fire-and-forget reduction (RED.ADD.F64), so no load round trip sits between TMEM reads
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int nn = n0 + cb * 32 + i;
                        if (nn < N) {
                            const double v = (double)__uint_as_float(hi[i]) + (double)__uint_as_float(lo[i]);
                            double *p = obase + (size_t)nn * ldo + m;
                            if (c == 0) *p = v; else atomicAdd(p, v);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA leaves while peers may still write to it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN) : "memory");
    }
}

// ---- operand preparation --------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(double x, __nv_bfloat16 &b0, __nv_bfloat16 &b1, __nv_bfloat16 &b2) {
    b0 = __double2bfloat16(x);
    const double r1 = x - (double)__bfloat162float(b0);
    b1 = __double2bfloat16(r1);
    const double r2 = r1 - (double)__bfloat162float(b1);
    b2 = __double2bfloat16(r2);
}

// planes[p][r][k] (row stride Kp) = bf16x3 split of x[r][k] (row stride ldx)
__global__ void split_rows_kernel(const double *__restrict__ x, long ldx, int rows, int K, int Kp,
                                  __nv_bfloat16 *__restrict__ planes) {
    const int r = blockIdx.y;
    const size_t ps = (size_t)rows * Kp;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Kp; k += gridDim.x * blockDim.x) {
        __nv_bfloat16 b0, b1, b2;
        split3(k < K ? x[(size_t)r * ldx + k] : 0.0, b0, b1, b2);
        const size_t o = (size_t)r * Kp + k;
        planes[o] = b0; planes[ps + o] = b1; planes[2 * ps + o] = b2;
    }
}

// planes[p][c][r] (row stride Rp) = bf16x3 split of x[r][c]: the transposed operand (32 x 32 tiles through smem)
__global__ void split_transpose_kernel(const double *__restrict__ x, int R, int C, int Rp,
                                       __nv_bfloat16 *__restrict__ planes) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? x[(size_t)r * C + c] : 0.0;
    }
    __syncthreads();
    const size_t ps = (size_t)C * Rp;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < Rp) {
            __nv_bfloat16 b0, b1, b2;
            split3(tile[threadIdx.x][i], b0, b1, b2);
            const size_t o = (size_t)c * Rp + r;
            planes[o] = b0; planes[ps + o] = b1; planes[2 * ps + o] = b2;
        }
    }
}

// out[b][m] = (c_in ? c_in[b][m] : 0) + sign * sum_s part[s][b][m]
__global__ void combine_kernel(const double *__restrict__ part, int slices, long slice_stride, const double *__restrict__ c_in,
                               double sign, int ld, double *__restrict__ out) {
    const int b = blockIdx.y;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < ld; m += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < slices; s++) acc += part[(size_t)s * slice_stride + (size_t)b * ld + m];
        const size_t o = (size_t)b * ld + m;
        out[o] = (c_in ? c_in[o] : 0.0) + sign * acc;
    }
}

// out[b] = sum_l secsq[b][l] (deterministic block sum); zero for inactive codewords
__global__ void dense_sumsq_kernel(const double *__restrict__ secsq, int L, const int *__restrict__ active, double *__restrict__ out) {
    __shared__ double red[40];
    const int b = blockIdx.x;
    double sq = 0.0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) sq += secsq[(size_t)b * L + l];
    const double v = block_sum(sq, red);
    if (threadIdx.x == 0) out[b] = v;
}

// ---- column-sharded A over NVLink peer memory (no NCCL in the loop) ------------------------------------------------
// Every rank owns a receive area [2 parities][world slots][B n + B] doubles that all peers have mapped.  One kernel
// folds the K slices of this rank's partial A beta, appends its |beta|^2 and PUSHES the result into slot `rank` of
// every peer's area with plain stores over NVLink (the transfer overlaps the fold tile by tile); the last CTA to
// finish publishes `epoch` in every peer's flag word with a system-scope release.  A one-warp kernel then waits until
// all `world` local flags have reached the epoch, and the residual kernel adds the slots in rank order -- the same
// order on every rank, so z and tau stay bit-identical everywhere and all ranks take the same early exit.  Areas
// alternate with the parity of the epoch: a rank can run at most one exchange ahead of the slowest reader.
struct P2pDev {
    double *slots[SB_P2P_MAX];
    unsigned long long *flags[SB_P2P_MAX];
};

__global__ void p2p_push_kernel(const double *__restrict__ part, int slices, long slice_stride, const double *__restrict__ tail,
                                int n, int B, P2pDev pd, int rank, int world, long area_off, long S,
                                unsigned long long epoch, unsigned int *__restrict__ done) {
    const int b = blockIdx.y;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < slices; s++) acc += part[(size_t)s * slice_stride + (size_t)b * n + m];
        for (int r = 0; r < world; r++) pd.slots[r][area_off + (size_t)rank * S + (size_t)b * n + m] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double t = tail ? tail[b] : 0.0;
        for (int r = 0; r < world; r++) pd.slots[r][area_off + (size_t)rank * S + (size_t)B * n + b] = t;
    }
    __threadfence_system();  // this thread's peer stores are performed before the CTA is counted as done
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(done, 1u) == total - 1) {  // last CTA: every CTA's stores are out
            *done = 0;
            __threadfence_system();
            for (int r = 0; r < world; r++)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pd.flags[r] + rank), "l"(epoch) : "memory");
        }
    }
}

// one thread per peer: wait for its flag; a peer that never arrives trips the timeout (err = 1) instead of hanging
__global__ void p2p_wait_kernel(const unsigned long long *flags, int world, unsigned long long epoch, long long timeout_cycles,
                                int *err) {
    if ((int)threadIdx.x >= world) return;
    const long long t0 = clock64();
    unsigned long long v = 0;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
        if (v >= epoch) return;
    } while (clock64() - t0 < timeout_cycles);
    *err = 1;
}

// out[b] = sum over the ranks (in rank order) of the |beta|^2 tails of an area
__global__ void p2p_tail_kernel(const double *__restrict__ area, int world, long S, long tail_off, int B, double *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0;
    for (int r = 0; r < world; r++) acc += area[(size_t)r * S + tail_off + b];
    out[b] = acc;
}

// ---- AMP iteration pieces (sparc_ldpc.py:203-220) -------------------------------------------------------------
// s = beta + A^T z (partials summed in slice order); beta <- sqrt(n P_l) softmax_section(s sqrt(n P_l) / tau^2);
// writes beta (fp64), its bf16x3 planes and sum(beta^2) per section.  One warp per section; inactive codewords
// (early stop) are left untouched.
__global__ void dense_denoise_kernel(const double *__restrict__ part, int slices, long slice_stride,
                                     const double *__restrict__ Pl, const double *__restrict__ tau2,
                                     const int *__restrict__ active, int L, int M, int n, int LMp, int B,
                                     double *__restrict__ beta, __nv_bfloat16 *__restrict__ planes,
                                     double *__restrict__ secsq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int b = blockIdx.y, l = blockIdx.x * wpb + warp;
    if (l >= L || !active[b]) return;
    const size_t LM = (size_t)L * M;
    const double rt = sqrt((double)n * Pl[l]), c2 = rt / tau2[b];
    double *bsec = beta + (size_t)b * LM + (size_t)l * M;
    const double *psec = part + (size_t)b * LM + (size_t)l * M;
    double m = -INFINITY;
    for (int j = lane; j < M; j += 32) {
        double s = bsec[j];
        for (int sl = 0; sl < slices; sl++) s += psec[(size_t)sl * slice_stride + j];
        m = fmax(m, s * c2);
    }
    for (int d = 16; d; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    double sum = 0.0;
    for (int j = lane; j < M; j += 32) {
        double s = bsec[j];
        for (int sl = 0; sl < slices; sl++) s += psec[(size_t)sl * slice_stride + j];
        sum += exp(s * c2 - m);
    }
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const double sc = rt / sum;
    double sq = 0.0;
    const size_t ps = (size_t)B * LMp;
    __nv_bfloat16 *pl = planes + (size_t)b * LMp + (size_t)l * M;
    for (int j = lane; j < M; j += 32) {
        double s = bsec[j];
        for (int sl = 0; sl < slices; sl++) s += psec[(size_t)sl * slice_stride + j];
        const double v = exp(s * c2 - m) * sc;
        bsec[j] = v;
        sq += v * v;
        __nv_bfloat16 b0, b1, b2;
        split3(v, b0, b1, b2);
        pl[j] = b0; pl[ps + j] = b1; pl[2 * ps + j] = b2;
    }
    for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    if (lane == 0) secsq[(size_t)b * L + l] = sq;
}

// One CTA per codeword.  mode 0 (start): z = y - (part ? sum_s part : 0).  mode 1 (iteration t):
// z <- y - A beta + (z / tau^2) (P - |beta|^2 / n)   (:220).  Then tau' = sqrt(|z|^2 / n) (:203) and the stop rule
// (:204): tau' == tau, or |tau' - tau| <= 2^-27 tau because operands rounded to 2^-27 cannot reach an exact fp64
// fixed point.  Writes z (fp64), its bf16x3 planes, tau^2 for the next iteration, and the per-codeword state.
__global__ void dense_residual_kernel(int mode, int t, const double *__restrict__ part, int slices, long slice_stride,
                                      const double *__restrict__ y, const double *__restrict__ secsq,
                                      const double *__restrict__ Pl, int L, int n, int np, int B, double *__restrict__ z,
                                      __nv_bfloat16 *__restrict__ planes, double *__restrict__ tau2, double *__restrict__ last_tau,
                                      int *__restrict__ active, int *__restrict__ iters, int *__restrict__ n_exec,
                                      unsigned *__restrict__ flags, double *__restrict__ tau2_trace, int T,
                                      int *__restrict__ n_active, const double *__restrict__ sumsq_ext, double P_ext) {
    __shared__ double red[40];
    const int b = blockIdx.x;
    if (mode == 1 && !active[b]) return;
    double coef = 0.0;
    if (mode == 1) {
        double sumsq, P;
        if (sumsq_ext) {  // column-sharded A: |beta|^2 and P summed over all ranks
            sumsq = sumsq_ext[b];
            P = P_ext;
        } else {
            double sq = 0.0, pw = 0.0;
            for (int l = threadIdx.x; l < L; l += blockDim.x) { sq += secsq[(size_t)b * L + l]; pw += Pl[l]; }
            sumsq = block_sum(sq, red);
            P = block_sum(pw, red);
        }
        coef = (P - sumsq / (double)n) / tau2[b];
    }
    const size_t ps = (size_t)B * np;
    double acc2 = 0.0;
    for (int k = threadIdx.x; k < np; k += blockDim.x) {
        double v = 0.0;
        if (k < n) {
            double x = 0.0;
            if (part) for (int s = 0; s < slices; s++) x += part[(size_t)s * slice_stride + (size_t)b * n + k];
            v = y[(size_t)b * n + k] - x;
            if (mode == 1) v += z[(size_t)b * n + k] * coef;
            z[(size_t)b * n + k] = v;
            acc2 += v * v;
        }
        __nv_bfloat16 b0, b1, b2;
        split3(v, b0, b1, b2);
        const size_t o = (size_t)b * np + k;
        planes[o] = b0; planes[ps + o] = b1; planes[2 * ps + o] = b2;
    }
    const double tau = sqrt(block_sum(acc2, red) / (double)n);
    if (threadIdx.x == 0) {
        const int tn = (mode == 0) ? 0 : t + 1;  // index of the iteration that would use this tau
        if (mode == 1) n_exec[b] += 1;
        const double lt = last_tau[b];
        if (mode == 1 && tn < T && (tau == lt || fabs(tau - lt) <= lt * 7.450580596923828e-09)) {
            active[b] = 0;
            iters[b] = tn;
            flags[b] |= SB_AMP_STOPPED;
            atomicSub(n_active, 1);
        } else {
            last_tau[b] = tau;
            tau2[b] = tau * tau;
            if (tau2_trace && tn < T) tau2_trace[(size_t)b * T + tn] = tau * tau;
        }
    }
}

__global__ void dense_init_state_kernel(int B, int T, int *active, int *iters, int *n_exec, unsigned *flags, double *last_tau,
                                        int *n_active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *n_active = B;
    if (b >= B) return;
    active[b] = 1;
    iters[b] = T > 0 ? T - 1 : 0;
    n_exec[b] = 0;
    flags[b] = 0;
    last_tau[b] = 0.0;
}

// ---- host side ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled g_encode = nullptr;

static int get_encode() {
    if (g_encode) return SB_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(SB_ECUDA, "cuTensorMapEncodeTiled is not available from the driver%s", "");
    g_encode = (PFN_cuTensorMapEncodeTiled)fn;
    return SB_OK;
}

// tensor map over bf16 planes [3][rows][Kp] with logical extents (K, rows, 3): boxes of (32 k, box_rows, 1 plane),
// 64-byte swizzle, out-of-range elements read as zero
static int make_map(CUtensorMap *tm, const void *planes, int rows, int K, int Kp, int box_rows) {
    if (get_encode() != SB_OK) return SB_ECUDA;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 3};
    cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)rows * Kp * 2};
    cuuint32_t box[3] = {GK, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(planes), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SB_ECUDA, "cuTensorMapEncodeTiled failed%s (%ld)", "", (long)r);
    return SB_OK;
}

static int pad8(int v) { return (v + 7) & ~7; }

}  // namespace sb

using namespace sb;

struct sb_dense {
    int n, LM, np, LMp;       // np / LMp = row strides of the planes (multiples of 8 values = 16 bytes)
    __nv_bfloat16 *A;         // [3][n][LMp]   rows of A       (A-side operand of A beta:  M = n,  K = LM)
    __nv_bfloat16 *At;        // [3][LM][np]   rows of A^T     (A-side operand of A^T z:   M = LM, K = n)
    CUtensorMap mapA, mapAt;
    int sms;
    // workspace, grown on demand
    int capB;
    __nv_bfloat16 *zpl, *bpl;  // [3][capB][np], [3][capB][LMp]
    double *part;              // max(slices_z * capB * LM, slices_b * capB * n)
    size_t part_elems;
    double *z, *tau2, *last_tau, *secsq;
    int *active, *n_active;
    unsigned int *p2p_done;  // [0] CTA counter of p2p_push_kernel, [1] timeout flag of p2p_wait_kernel
    int secsq_L;
};

static void dense_free_ws(sb_dense *d) {
    cudaFree(d->zpl); cudaFree(d->bpl); cudaFree(d->part); cudaFree(d->z); cudaFree(d->tau2); cudaFree(d->last_tau);
    cudaFree(d->secsq); cudaFree(d->active); cudaFree(d->n_active); cudaFree(d->p2p_done);
    d->p2p_done = nullptr;
    d->zpl = d->bpl = nullptr; d->part = d->z = d->tau2 = d->last_tau = d->secsq = nullptr; d->active = d->n_active = nullptr;
    d->capB = 0; d->part_elems = 0; d->secsq_L = 0;
}

// K slices: enough CTAs to fill the GPU (the A beta product has only n / 128 row tiles), never more than k-blocks
static void plan_slices(const sb_dense *d, int Mrows, int K, int B, int BN, int *slices, int *kb_per_slice) {
    const int nkb = (K + GK - 1) / GK;
    const int tiles = ((Mrows + GM - 1) / GM) * ((B + BN - 1) / BN);
    int s = (2 * d->sms + tiles - 1) / tiles;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    if (s > nkb) s = nkb;
    int per = (nkb + s - 1) / s;
    const char *env = knob("SB_DENSE_CHUNK_KB");
    const int chunk = env ? atoi(env) : 32;
    if (per > chunk) per = ((per + chunk - 1) / chunk) * chunk;  // whole chunks per slice
    *kb_per_slice = per;
    *slices = (nkb + per - 1) / per;
}

static int pick_bn(int B) {
    const char *env = knob("SB_DENSE_BN");
    if (env) return atoi(env) == 256 ? 256 : 128;
    return B > 128 ? 256 : 128;
}

static int dense_reserve(sb_dense *d, int B, int L) {
    size_t need_part = 0;
    {
        int s1, p1, s2, p2;
        const int BN = pick_bn(B);
        plan_slices(d, d->LM, d->n, B, BN, &s1, &p1);
        plan_slices(d, d->n, d->LM, B, BN, &s2, &p2);
        need_part = std::max((size_t)s1 * B * d->LM, (size_t)s2 * B * d->n);
    }
    if (B <= d->capB && need_part <= d->part_elems && L <= d->secsq_L) return SB_OK;
    dense_free_ws(d);
    SB_CUDA(cudaMalloc(&d->zpl, sizeof(__nv_bfloat16) * 3 * (size_t)B * d->np));
    SB_CUDA(cudaMalloc(&d->bpl, sizeof(__nv_bfloat16) * 3 * (size_t)B * d->LMp));
    SB_CUDA(cudaMalloc(&d->part, sizeof(double) * need_part));
    SB_CUDA(cudaMalloc(&d->z, sizeof(double) * (size_t)B * d->n));
    SB_CUDA(cudaMalloc(&d->tau2, sizeof(double) * B));
    SB_CUDA(cudaMalloc(&d->last_tau, sizeof(double) * B));
    SB_CUDA(cudaMalloc(&d->secsq, sizeof(double) * (size_t)B * (L > 0 ? L : 1)));
    SB_CUDA(cudaMalloc(&d->active, sizeof(int) * B));
    SB_CUDA(cudaMalloc(&d->n_active, sizeof(int)));
    SB_CUDA(cudaMalloc(&d->p2p_done, 2 * sizeof(unsigned int)));
    d->capB = B; d->part_elems = need_part; d->secsq_L = L > 0 ? L : 1;
    return SB_OK;
}

// part[s][b][m] = sum over slice s of Aop[m][k] * xplanes[b][k];  transpose = 0: Aop = A (M = n, K = LM), 1: A^T
template <int BN, int STAGES>
static int launch_gemm(dim3 grid, int CL, cudaStream_t st, const CUtensorMap &mapA, const CUtensorMap &mapB, double *part,
                       int Mrows, int B, long sstride, int nkb, int per, int chunk) {
    using S = GemmSmem<BN, STAGES>;
    auto kern = gemm_bf16x3_kernel<BN, STAGES>;
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(192, 1, 1);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SB_CUDA(cudaLaunchKernelEx(&cfg, kern, mapA, mapB, part, Mrows, B, Mrows, sstride, nkb, per, chunk, CL));
    return SB_OK;
}

static int dense_gemm_launch(const sb_dense *d, int transpose, const __nv_bfloat16 *xpl, int B, double *part,
                             int *slices_out, long *slice_stride_out, cudaStream_t st) {
    const int Mrows = transpose ? d->LM : d->n, K = transpose ? d->n : d->LM, Kp = transpose ? d->np : d->LMp;
    const int BN = pick_bn(B);
    const int mtiles = (Mrows + GM - 1) / GM;
    // cluster size: the codeword-side tile is fetched once per cluster (SB_DENSE_CLUSTER = 1 | 2 | 4).  Measured at n = 4608,
    // LM = 65536, B = 256 (tools/r2_dense.sh): A beta (36 row tiles) 0.89 ms with 2, 1.17 ms with 4; A^T z (512 row tiles)
    // 0.99 ms with 2, 0.95 ms with 4
    int CL = transpose ? 4 : 2;
    if (const char *env = knob("SB_DENSE_CLUSTER")) CL = atoi(env);
    if (CL != 1 && CL != 2 && CL != 4) CL = 2;
    while (CL > 1 && mtiles < CL) CL /= 2;
    CUtensorMap mapB;
    int rc = make_map(&mapB, xpl, B, K, Kp, BN / CL);
    if (rc != SB_OK) return rc;
    int slices, per;
    plan_slices(d, Mrows, K, B, BN, &slices, &per);
    const int nkb = (K + GK - 1) / GK;
    const char *env = knob("SB_DENSE_CHUNK_KB");
    const int chunk = env ? atoi(env) : 32;
    const long sstride = (long)B * Mrows;
    // row tiles padded to whole clusters: the extra CTAs read zeros (TMA out-of-range fill) and store nothing
    dim3 grid(((mtiles + CL - 1) / CL) * CL, (B + BN - 1) / BN, slices);
    const CUtensorMap &mapA = transpose ? d->mapAt : d->mapA;
    if (BN == 256) rc = launch_gemm<256, 3>(grid, CL, st, mapA, mapB, part, Mrows, B, sstride, nkb, per, chunk);
    else rc = launch_gemm<128, 4>(grid, CL, st, mapA, mapB, part, Mrows, B, sstride, nkb, per, chunk);
    if (rc != SB_OK) return rc;
    SB_LAUNCHED();
    *slices_out = slices;
    *slice_stride_out = sstride;
    return SB_OK;
}

extern "C" int sb_dense_create(const double *A_dev, int n, int LM, sb_dense **out) {
    if (!A_dev || !out || n <= 0 || LM <= 0) return fail(SB_EINVAL, "sb_dense_create: bad argument%s", "");
    int dev = 0, major = 0, sms = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) return fail(SB_ECUDA, "sb_dense_create: the tcgen05 GEMM needs an sm_100 device%s (found major %ld)", "", major);
    sb_dense *d = new sb_dense();
    memset(d, 0, sizeof(*d));
    d->n = n; d->LM = LM; d->np = pad8(n); d->LMp = pad8(LM); d->sms = sms;
    cudaError_t e1 = cudaMalloc(&d->A, sizeof(__nv_bfloat16) * 3 * (size_t)n * d->LMp);
    cudaError_t e2 = cudaMalloc(&d->At, sizeof(__nv_bfloat16) * 3 * (size_t)LM * d->np);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cudaFree(d->A); cudaFree(d->At); delete d;
        return fail(SB_ENOMEM, "sb_dense_create: cudaMalloc failed%s", "");
    }
    {
        dim3 g((d->LMp + 255) / 256 < 1024 ? (d->LMp + 255) / 256 : 1024, n);
        split_rows_kernel<<<g, 256>>>(A_dev, LM, n, LM, d->LMp, d->A);
        g_launches.fetch_add(1);
        dim3 gt((LM + 31) / 32, (d->np + 31) / 32);
        split_transpose_kernel<<<gt, dim3(32, 8)>>>(A_dev, n, LM, d->np, d->At);
        g_launches.fetch_add(1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(d->A); cudaFree(d->At); delete d; return fail(SB_ECUDA, "sb_dense_create: %s", cudaGetErrorString(e)); }
    int rc = make_map(&d->mapA, d->A, n, LM, d->LMp, GM);
    if (rc == SB_OK) rc = make_map(&d->mapAt, d->At, LM, n, d->np, GM);
    if (rc != SB_OK) { cudaFree(d->A); cudaFree(d->At); delete d; return rc; }
    {   // load the kernels of the peer-memory exchange now: a lazy module load at their first launch synchronises the
        // context, which deadlocks if another rank of the SAME process (threads) is already spinning in p2p_wait_kernel
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, p2p_push_kernel);
        cudaFuncGetAttributes(&fa, p2p_wait_kernel);
        cudaFuncGetAttributes(&fa, p2p_tail_kernel);
        cudaFuncGetAttributes(&fa, dense_sumsq_kernel);
        cudaGetLastError();
    }
    *out = d;
    return SB_OK;
}

extern "C" void sb_dense_destroy(sb_dense *d) {
    if (!d) return;
    dense_free_ws(d);
    cudaFree(d->A);
    cudaFree(d->At);
    delete d;
}

// out[b] = A x[b] (transpose = 0: x [B][LM] -> out [B][n]) or A^T x[b] (transpose = 1: x [B][n] -> out [B][LM])
extern "C" int sb_dense_apply_batch(sb_dense *d, int transpose, const double *x, int B, double *out, void *stream) {
    if (!d || !x || !out || B < 0) return fail(SB_EINVAL, "sb_dense_apply_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dense_reserve(d, B, 1);
    if (rc != SB_OK) return rc;
    const int K = transpose ? d->n : d->LM, Kp = transpose ? d->np : d->LMp, Mrows = transpose ? d->LM : d->n;
    __nv_bfloat16 *pl = transpose ? d->zpl : d->bpl;
    dim3 g((Kp + 255) / 256 < 1024 ? (Kp + 255) / 256 : 1024, B);
    split_rows_kernel<<<g, 256, 0, st>>>(x, K, B, K, Kp, pl);
    SB_LAUNCHED();
    int slices; long ss;
    rc = dense_gemm_launch(d, transpose, pl, B, d->part, &slices, &ss, st);
    if (rc != SB_OK) return rc;
    dim3 gc((Mrows + 255) / 256 < 1024 ? (Mrows + 255) / 256 : 1024, B);
    combine_kernel<<<gc, 256, 0, st>>>(d->part, slices, ss, nullptr, 1.0, Mrows, out);
    SB_LAUNCHED();
    return SB_OK;
}

// Batched AMP decode with the dense operator (sparc_ldpc.py:189-222).  All pointers are device pointers:
// y [B][n], Pl [L], beta0 [B][L*M] or NULL, beta [B][L*M] out, iters / n_exec / flags [B] out,
// tau2_trace [B][T] or NULL.  The loop stops early once every codeword has stopped (checked every 4 iterations).
//
// Column-sharded mode (allreduce != NULL): this handle holds the columns of L local sections of a larger matrix;
// Pl / beta0 / beta refer to the local sections, P_total = sum of Pl over ALL sections, z and tau^2 are replicated
// on every rank.  A^T z, the softmax and |beta|^2 are local; the partial A beta of every rank plus its |beta|^2
// ([B][n] + [B] doubles in xbuf) are summed over the ranks by the caller's `allreduce` once per iteration.
static int dense_amp_impl(sb_dense *d, const double *y, const double *Pl, double P_total, const double *beta0, int L, int M,
                          int B, int T, double *beta, int *iters, int *n_exec, unsigned *flags, double *tau2_trace,
                          sb_allreduce_fn allreduce, void *ctx, double *xbuf, cudaStream_t st, const sb_p2p *px = nullptr,
                          unsigned long long *epoch_io = nullptr) {
    int rc = dense_reserve(d, B, L);
    if (rc != SB_OK) return rc;
    const int n = d->n, LM = d->LM;
    dense_init_state_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, T, d->active, iters, n_exec, flags, d->last_tau, d->n_active);
    SB_LAUNCHED();
    if (tau2_trace) SB_CUDA(cudaMemsetAsync(tau2_trace, 0xFF, sizeof(double) * (size_t)B * T, st));  // NaN = not executed
    int slices = 0; long ss = 0;
    // after a GEMM of partial A beta: (sharded) fold the K slices into xbuf, append |beta|^2, sum over the ranks
    // peer-memory exchange state (px != NULL): S doubles per slot, areas alternate with the epoch's parity
    // (the slot / parity strides are the ALLOCATED slot size, not this call's B*n + B: a call with a smaller batch
    // must not put its parity-1 area over memory that a slower peer may still be reading from the previous call)
    const long S = px ? px->slot_doubles : (long)B * n + B;
    if (px && (px->slot_doubles < (long)B * n + B || px->world < 1 || px->world > SB_P2P_MAX))
        return fail(SB_EINVAL, "sb_dense_amp_batch_p2p: the peer areas hold %s%ld doubles per slot, fewer than B*n + B", "", px->slot_doubles);
    P2pDev pd;
    unsigned long long epoch = epoch_io ? *epoch_io : 0;
    struct EpochGuard {  // the host counter follows the flags already published, on every exit path
        unsigned long long *io, *cur;
        ~EpochGuard() { if (io) *io = *cur; }
    } epoch_guard{epoch_io, &epoch};
    const double *p2p_area = nullptr;
    if (px) {
        for (int r = 0; r < px->world; r++) { pd.slots[r] = px->slots[r]; pd.flags[r] = px->flags[r]; }
        SB_CUDA(cudaMemsetAsync(d->p2p_done, 0, sizeof(unsigned int) + sizeof(int), st));
    }
    auto exchange_p2p = [&](bool with_sumsq) -> int {
        epoch++;
        const long area_off = (long)(epoch & 1ull) * px->world * S;
        if (with_sumsq) {
            dense_sumsq_kernel<<<B, 128, 0, st>>>(d->secsq, L, d->active, xbuf + (size_t)B * n);
            SB_LAUNCHED();
        }
        dim3 gc((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024, B);
        p2p_push_kernel<<<gc, 256, 0, st>>>(d->part, slices, ss, with_sumsq ? xbuf + (size_t)B * n : nullptr, n, B, pd,
                                            px->rank, px->world, area_off, S, epoch, d->p2p_done);
        SB_LAUNCHED();
        p2p_wait_kernel<<<1, 32, 0, st>>>(px->flags[px->rank], px->world, epoch, (long long)px->timeout_ms * 2000000ll,
                                          reinterpret_cast<int *>(d->p2p_done + 1));
        SB_LAUNCHED();
        p2p_area = px->slots[px->rank] + area_off;
        p2p_tail_kernel<<<(B + 127) / 128, 128, 0, st>>>(p2p_area, px->world, S, (long)B * n, B, xbuf + (size_t)B * n);
        SB_LAUNCHED();
        return SB_OK;
    };
    auto exchange = [&](bool with_sumsq) -> int {
        if (px) return exchange_p2p(with_sumsq);
        if (!allreduce) return SB_OK;
        dim3 gc((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024, B);
        combine_kernel<<<gc, 256, 0, st>>>(d->part, slices, ss, nullptr, 1.0, n, xbuf);
        SB_LAUNCHED();
        if (with_sumsq) {
            dense_sumsq_kernel<<<B, 128, 0, st>>>(d->secsq, L, d->active, xbuf + (size_t)B * n);
            SB_LAUNCHED();
        } else {
            SB_CUDA(cudaMemsetAsync(xbuf + (size_t)B * n, 0, sizeof(double) * B, st));
        }
        const int r = allreduce(ctx, xbuf, (long)B * n + B, (void *)st);
        if (r != 0) return fail(SB_ECUDA, "sb_dense_amp_batch_sharded: the allreduce callback failed%s (%ld)", "", (long)r);
        return SB_OK;
    };
    const bool ext = allreduce || px;
    const double *xsrc = allreduce ? xbuf : d->part;
    const double *sq_ext = ext ? xbuf + (size_t)B * n : nullptr;
    if (beta0) {  // z = y - A beta0   (:197-198)
        if (beta0 != beta) SB_CUDA(cudaMemcpyAsync(beta, beta0, sizeof(double) * (size_t)B * LM, cudaMemcpyDeviceToDevice, st));
        dim3 g((d->LMp + 255) / 256 < 1024 ? (d->LMp + 255) / 256 : 1024, B);
        split_rows_kernel<<<g, 256, 0, st>>>(beta, LM, B, LM, d->LMp, d->bpl);
        SB_LAUNCHED();
        rc = dense_gemm_launch(d, 0, d->bpl, B, d->part, &slices, &ss, st);
        if (rc != SB_OK) return rc;
        rc = exchange(false);
        if (rc != SB_OK) return rc;
    } else {
        SB_CUDA(cudaMemsetAsync(beta, 0, sizeof(double) * (size_t)B * LM, st));
    }
    // partial sums the residual kernel adds up: K slices (one GPU), the all-reduced buffer, or the world's slots
    auto res_src = [&]() { return px ? p2p_area : xsrc; };
    auto res_cnt = [&]() { return px ? px->world : (allreduce ? 1 : slices); };
    auto res_str = [&]() { return px ? S : (allreduce ? (long)B * n : ss); };
    dense_residual_kernel<<<B, 256, 0, st>>>(0, 0, beta0 ? res_src() : nullptr, res_cnt(), res_str(), y,
                                             d->secsq, Pl, L, n, d->np, B, d->z, d->zpl, d->tau2, d->last_tau, d->active, iters,
                                             n_exec, flags, tau2_trace, T, d->n_active, sq_ext, P_total);
    SB_LAUNCHED();
    for (int t = 0; t < T; t++) {
        rc = dense_gemm_launch(d, 1, d->zpl, B, d->part, &slices, &ss, st);  // A^T z
        if (rc != SB_OK) return rc;
        dim3 gd((L + 7) / 8, B);
        dense_denoise_kernel<<<gd, 256, 0, st>>>(d->part, slices, ss, Pl, d->tau2, d->active, L, M, n, d->LMp, B, beta, d->bpl,
                                                 d->secsq);
        SB_LAUNCHED();
        rc = dense_gemm_launch(d, 0, d->bpl, B, d->part, &slices, &ss, st);  // A beta
        if (rc != SB_OK) return rc;
        rc = exchange(true);
        if (rc != SB_OK) return rc;
        dense_residual_kernel<<<B, 256, 0, st>>>(1, t, res_src(), res_cnt(), res_str(), y, d->secsq, Pl,
                                                 L, n, d->np, B, d->z, d->zpl, d->tau2, d->last_tau, d->active, iters, n_exec,
                                                 flags, tau2_trace, T, d->n_active, sq_ext, P_total);
        SB_LAUNCHED();
        if ((t & 3) == 3 && t + 1 < T) {  // z and tau are replicated bit for bit, so every rank takes the same exit
            int na = 0;
            SB_CUDA(cudaMemcpyAsync(&na, d->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            if (na <= 0) break;
        }
    }
    if (px) {
        int perr = 0;
        SB_CUDA(cudaMemcpyAsync(&perr, d->p2p_done + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        if (perr) return fail(SB_ECUDA, "sb_dense_amp_batch_p2p: a peer did not arrive within the timeout%s", "");
    }
    return SB_OK;
}

extern "C" int sb_dense_amp_batch(sb_dense *d, const double *y, const double *Pl, const double *beta0, int L, int M, int B,
                                  int T, double *beta, int *iters, int *n_exec, unsigned *flags, double *tau2_trace,
                                  void *stream) {
    if (!d || !y || !Pl || !beta || !iters || !n_exec || !flags || B < 0 || T < 0 || L <= 0 || M <= 0)
        return fail(SB_EINVAL, "sb_dense_amp_batch: bad argument%s", "");
    if ((long)L * M != d->LM) return fail(SB_EINVAL, "sb_dense_amp_batch: L*M does not match the matrix%s (%ld)", "", (long)L * M);
    if (B == 0) return SB_OK;
    return dense_amp_impl(d, y, Pl, 0.0, beta0, L, M, B, T, beta, iters, n_exec, flags, tau2_trace, nullptr, nullptr, nullptr,
                          (cudaStream_t)stream);
}

// lets kernels of the current device store into / load from `peer_device`'s memory (IPC-mapped receive areas)
extern "C" int sb_enable_peer_access(int peer_device) {
    int cur = -1;
    SB_CUDA(cudaGetDevice(&cur));
    if (cur == peer_device) return SB_OK;
    int can = 0;
    SB_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    if (!can) return fail(SB_ECUDA, "sb_enable_peer_access: no peer access to device%s %ld", "", (long)peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SB_OK; }
    SB_CUDA(e);
    return SB_OK;
}

// Receive areas of the peer-memory exchange: plain cudaMalloc blocks (zeroed) exported / imported with CUDA IPC.  The
// import runs with the ACCESSING device current and cudaIpcMemLazyEnablePeerAccess, which is what maps the peer's
// memory into this device's address space (a handle opened under the owner's device is not reachable from here).
extern "C" int sb_p2p_alloc(long bytes, void **dev_ptr, unsigned char *handle64) {
    if (!dev_ptr || !handle64 || bytes <= 0) return fail(SB_EINVAL, "sb_p2p_alloc: bad argument%s", "");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void *p = nullptr;
    SB_CUDA(cudaMalloc(&p, (size_t)bytes));
    SB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); SB_CUDA(e); }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return SB_OK;
}
extern "C" int sb_p2p_open(const unsigned char *handle64, void **dev_ptr) {
    if (!dev_ptr || !handle64) return fail(SB_EINVAL, "sb_p2p_open: bad argument%s", "");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SB_OK;
}
extern "C" int sb_p2p_close(void *dev_ptr) {
    if (dev_ptr) SB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return SB_OK;
}
extern "C" int sb_p2p_free(void *dev_ptr) {
    if (dev_ptr) SB_CUDA(cudaFree(dev_ptr));
    return SB_OK;
}

extern "C" int sb_dense_amp_batch_p2p(sb_dense *d, const double *y, const double *Pl_local, double P_total,
                                      const double *beta0_local, int L_local, int M, int B, int T, double *beta_local,
                                      int *iters, int *n_exec, unsigned *flags, double *tau2_trace, double *xbuf,
                                      const sb_p2p *px, unsigned long long *epoch_io, void *stream) {
    if (!d || !y || !Pl_local || !beta_local || !iters || !n_exec || !flags || !xbuf || !px || !epoch_io || B < 0 || T < 0 ||
        L_local <= 0 || M <= 0 || px->world < 1 || px->world > SB_P2P_MAX || px->rank < 0 || px->rank >= px->world)
        return fail(SB_EINVAL, "sb_dense_amp_batch_p2p: bad argument%s", "");
    for (int r = 0; r < px->world; r++)
        if (!px->slots[r] || !px->flags[r]) return fail(SB_EINVAL, "sb_dense_amp_batch_p2p: peer %s%ld not mapped", "", (long)r);
    if ((long)L_local * M != d->LM)
        return fail(SB_EINVAL, "sb_dense_amp_batch_p2p: L_local*M does not match the matrix%s (%ld)", "", (long)L_local * M);
    if (B == 0) return SB_OK;
    return dense_amp_impl(d, y, Pl_local, P_total, beta0_local, L_local, M, B, T, beta_local, iters, n_exec, flags, tau2_trace,
                          nullptr, nullptr, xbuf, (cudaStream_t)stream, px, epoch_io);
}

extern "C" int sb_dense_amp_batch_sharded(sb_dense *d, const double *y, const double *Pl_local, double P_total,
                                          const double *beta0_local, int L_local, int M, int B, int T, double *beta_local,
                                          int *iters, int *n_exec, unsigned *flags, double *tau2_trace, double *xbuf,
                                          sb_allreduce_fn allreduce, void *ctx, void *stream) {
    if (!d || !y || !Pl_local || !beta_local || !iters || !n_exec || !flags || !xbuf || !allreduce || B < 0 || T < 0 ||
        L_local <= 0 || M <= 0)
        return fail(SB_EINVAL, "sb_dense_amp_batch_sharded: bad argument%s", "");
    if ((long)L_local * M != d->LM)
        return fail(SB_EINVAL, "sb_dense_amp_batch_sharded: L_local*M does not match the matrix%s (%ld)", "", (long)L_local * M);
    if (B == 0) return SB_OK;
    return dense_amp_impl(d, y, Pl_local, P_total, beta0_local, L_local, M, B, T, beta_local, iters, n_exec, flags, tau2_trace,
                          allreduce, ctx, xbuf, (cudaStream_t)stream);
}
