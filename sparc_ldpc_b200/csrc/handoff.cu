// handoff.cu -- AMP <-> LDPC soft handoff, decisions, peeling and EXIT histograms (sm_100a, fp64).
//
// Replaces the pure-Python loops of the reference: sp2bp (ldpc/sparc_ldpc.py:257-281) + LLR conversion
// (:667-669), bp2sp (:283-314) + prior scaling (:685-696), argmax / bits2indices / BER counting
// (:317-356, :640-650), hard_initialisation's threshold test (ldpc/amp_exit.py:85-106) and the two
// np.histogram calls of hist_E (amp_exit.py:303-305).  All are one pass over beta or over L*logM bits:
// HBM-bound by construction.  Summation orders follow the reference (sequential ascending-j adds in sp2bp,
// sequential MSB-first products and numpy's pairwise section sum in bp2sp), so given identical inputs the
// outputs are bit-identical up to CUDA-vs-numpy exp/log ulps.  This file is compiled with -fmad=false.
#include <cfloat>

#include "common.cuh"

namespace sb {

__device__ __forceinline__ double nan_to_num(double x) {  // np.nan_to_num (sparc_ldpc.py:669)
    if (x != x) return 0.0;
    if (x == INFINITY) return DBL_MAX;
    if (x == -INFINITY) return -DBL_MAX;
    return x;
}

// One warp per section entry.  smem: M doubles per warp.
__global__ void sp2bp_llr_kernel(const double *__restrict__ beta, long beta_stride, int beta_first,
                                 const int *__restrict__ sections, const int *__restrict__ nsec, int L_stride,
                                 int first_sec, int out_first, int count, int M, int logM, int n,
                                 const double *__restrict__ Pl, double *__restrict__ p_out, double *__restrict__ llr,
                                 long out_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double *s = reinterpret_cast<double *>(smem_raw) + (size_t)warp * M;
    const int b = blockIdx.y;
    const int cnt = sections ? nsec[b] : count;
    const int i = blockIdx.x * wpb + warp;
    if (i >= cnt) return;
    const int sec = sections ? sections[(size_t)b * L_stride + i] : first_sec + i;
    const int osec = sections ? sec : out_first + i;
    const double scale = sqrt((double)n * Pl[sec]);  // np.sqrt(n*np.repeat(Pl, M))   (sparc_ldpc.py:657)
    const double *src = beta + (size_t)b * beta_stride + (size_t)(beta_first + i) * M;
    for (int j = lane; j < M; j += 32) s[j] = src[j] / scale;
    __syncwarp();
    for (int logi = lane; logi < logM; logi += 32) {  // logM <= 10 < 32: one bit per lane
        double acc = 0.0;
        const int lowmask = (1 << logi) - 1;
        for (int t = 0; t < M / 2; t++) {  // ascending j over the indices whose bit `logi` is set (:276-280)
            const int j = ((t >> logi) << (logi + 1)) | (1 << logi) | (t & lowmask);
            acc = acc + s[j];
        }
        const size_t pos = (size_t)b * out_stride + (size_t)osec * logM + (logM - logi - 1);
        if (p_out) p_out[pos] = acc;
        llr[pos] = nan_to_num(log(1 - acc) - log(acc));
    }
}

// The same map for M >= 64 with every lane busy.  The sequential ascending-j chain of one (section, bit) pair cannot
// be split without changing the rounding, so the parallelism comes from the pairs: a CTA stages R sections in
// shared memory (row stride M + 1 doubles: the rows of a bit group hit distinct bank pairs), then R lanes run
// the chains of one bit of the R sections, 32 / R bits per warp.  The
// one-warp-per-section kernel above keeps 9 of 32 lanes busy in its chain phase (1.2 TB/s at M = 512).
template <int R>  // sections per CTA (8 | 16): lanes r + R g of warp w run the chains of bit w (32 / R) + g
__global__ void sp2bp_llr_kernel16(const double *__restrict__ beta, long beta_stride, int beta_first,
                                   const int *__restrict__ sections, const int *__restrict__ nsec, int L_stride,
                                   int first_sec, int out_first, int count, int M, int logM, int n,
                                   const double *__restrict__ Pl, double *__restrict__ p_out, double *__restrict__ llr,
                                   long out_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s = reinterpret_cast<double *>(smem_raw);
    constexpr int G = 32 / R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwl = blockDim.x >> 5, nw = (logM + G - 1) / G;
    const int b = blockIdx.y;
    const int cnt = sections ? nsec[b] : count;
    const int i0 = blockIdx.x * R;
    if (i0 >= cnt) return;
    const int rows = min(R, cnt - i0), rs = M + 1;
    for (int r = warp; r < rows; r += nwl) {  // all 8 warps load (enough bytes in flight for HBM); nw of them run chains
        const int sec = sections ? sections[(size_t)b * L_stride + i0 + r] : first_sec + i0 + r;
        const double scale = sqrt((double)n * Pl[sec]);  // np.sqrt(n*np.repeat(Pl, M))   (sparc_ldpc.py:657)
        const double *src = beta + (size_t)b * beta_stride + (size_t)(beta_first + i0 + r) * M;
        for (int j0 = 0; j0 < M; j0 += 512) {  // 16 loads in flight per lane, then the divisions (whose slow-path
            double v[16];                        // branches would otherwise serialise load -> divide -> store)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int j = j0 + 32 * k + lane;
                v[k] = (j < M) ? src[j] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int j = j0 + 32 * k + lane;
                if (j < M) s[r * rs + j] = v[k] / scale;
            }
        }
    }
    __syncthreads();
    const int r = lane % R, logi = warp * G + lane / R;
    if (warp >= nw || r >= rows || logi >= logM) return;
    const double *row = s + r * rs;
    double acc = 0.0;
    const int lowmask = (1 << logi) - 1;
#pragma unroll 8
    for (int t = 0; t < M / 2; t++) {  // ascending j over the indices whose bit `logi` is set (:276-280)
        const int j = ((t >> logi) << (logi + 1)) | (1 << logi) | (t & lowmask);
        acc = acc + row[j];
    }
    const int sec = sections ? sections[(size_t)b * L_stride + i0 + r] : first_sec + i0 + r;
    const int osec = sections ? sec : out_first + i0 + r;
    const size_t pos = (size_t)b * out_stride + (size_t)osec * logM + (logM - logi - 1);
    if (p_out) p_out[pos] = acc;
    llr[pos] = nan_to_num(log(1 - acc) - log(acc));
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src) of a[0..n), n a power of two <= 128,
// executed by one thread.
__device__ __forceinline__ double np_block_sum(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    for (int i = 8; i < n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    }
    return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
}

// One warp per section.  smem: M doubles per warp (+ 16 for block sums).
__global__ void bp2sp_prior_kernel(const double *__restrict__ app, int ls, const double *__restrict__ beta_prev, int L,
                                   int M, int logM, int n, const double *__restrict__ Pl, int scale_by_power,
                                   int input_is_prob, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double *s = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (M + 16);
    double *bs = s + M;
    const int b = blockIdx.y;
    const int l = blockIdx.x * wpb + warp;
    if (l >= L) return;
    const double scale = sqrt((double)n * Pl[l]);
    double *dst = out + ((size_t)b * L + l) * M;
    if (l < L - ls) {  // unprotected section: posterior = beta/scale, prior = posterior*scale (:657, :696)
        const double *src = beta_prev + ((size_t)b * L + l) * M;
        for (int j = lane; j < M; j += 32) {
            const double post = src[j] / scale;
            dst[j] = scale_by_power ? post * scale : post;
        }
        return;
    }
    // bitwise = 1/(1+exp(app))  (:685), kept one per lane
    double v = 0.0;
    if (lane < logM) {
        const double in = app[(size_t)b * ls * logM + (size_t)(l - (L - ls)) * logM + lane];
        v = input_is_prob ? in : 1 / (1 + exp(in));
    }
    for (int m0 = 0; m0 < M; m0 += 32) {  // np.prod over the logM bits, MSB first (:301-312)
        const int m = m0 + lane;
        double prod = 1.0;
        for (int jb = 0; jb < logM; jb++) {
            const double vj = __shfl_sync(0xffffffffu, v, jb);
            const int bit = (m >> (logM - 1 - jb)) & 1;
            prod = prod * (bit ? vj : (1 - vj));
        }
        if (m < M) s[m] = prod;
    }
    __syncwarp();
    // np.sum of the section (:313): leaves of <= 128 elements, then a balanced tree
    const int leaf = M < 128 ? M : 128, nleaf = M / leaf;
    for (int i = lane; i < nleaf; i += 32) bs[i] = np_block_sum(s + i * leaf, leaf);
    __syncwarp();
    double tot;
    {
        double t8[8];
        for (int i = 0; i < nleaf; i++) t8[i] = bs[i];
        for (int w2 = nleaf; w2 > 1; w2 >>= 1)
            for (int i = 0; i < w2 / 2; i++) t8[i] = t8[2 * i] + t8[2 * i + 1];
        tot = t8[0];
    }
    for (int m = lane; m < M; m += 32) {
        const double post = s[m] / tot;
        dst[m] = scale_by_power ? post * scale : post;
    }
}

// M = 512 with every lane busy in the section sum.  numpy's pairwise sum of 512 values = 4 leaves of 128, each
// leaf 8 interleaved accumulators of 16 sequential terms (r[j] += a[i + j], i = 8, 16, ...), then
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) per leaf and (l0+l1)+(l2+l3): lane 8 f + j runs accumulator j of leaf f
// (32 chains of 16 terms; the leaves are staggered by 0, 8, 8, 16 doubles so that a half-warp's 16 lanes hit 16
// distinct bank pairs), the two trees are xor-shuffles 1, 2, 4 and 8, 16.  Same operands in the same order as
// bp2sp_prior_kernel, hence the same bits; the bit posteriors are shuffled once per section, not once per product.
__global__ void bp2sp_prior_kernel512(const double *__restrict__ app, int ls, const double *__restrict__ beta_prev, int L,
                                      int n, const double *__restrict__ Pl, int scale_by_power, int input_is_prob,
                                      double *__restrict__ out) {
    constexpr int M = 512, logM = 9;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double *s = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (M + 16);
    const int b = blockIdx.y;
    const int l = blockIdx.x * wpb + warp;
    if (l >= L) return;
    const double scale = sqrt((double)n * Pl[l]);
    double *dst = out + ((size_t)b * L + l) * M;
    if (l < L - ls) {  // unprotected section: posterior = beta/scale, prior = posterior*scale (:657, :696)
        const double *src = beta_prev + ((size_t)b * L + l) * M;
#pragma unroll 4
        for (int j = lane; j < M; j += 32) {
            const double post = src[j] / scale;
            dst[j] = scale_by_power ? post * scale : post;
        }
        return;
    }
    double v = 0.0;  // bitwise = 1/(1+exp(app))  (:685), one per lane
    if (lane < logM) {
        const double in = app[(size_t)b * ls * logM + (size_t)(l - (L - ls)) * logM + lane];
        v = input_is_prob ? in : 1 / (1 + exp(in));
    }
    double vb[logM], nb[logM];
#pragma unroll
    for (int jb = 0; jb < logM; jb++) {
        vb[jb] = __shfl_sync(0xffffffffu, v, jb);
        nb[jb] = 1 - vb[jb];
    }
    double x[M / 32];
#pragma unroll
    for (int k = 0; k < M / 32; k++) {  // np.prod over the logM bits, MSB first (:301-312)
        const int m = 32 * k + lane;
        double prod = 1.0;
#pragma unroll
        for (int jb = 0; jb < logM; jb++) prod = prod * (((m >> (logM - 1 - jb)) & 1) ? vb[jb] : nb[jb]);
        x[k] = prod;
        s[m + 8 * (((m >> 7) + 1) >> 1)] = prod;  // leaf f starts 8 ((f + 1) >> 1) doubles further: bank pairs 0, 8, 8, 0
    }
    __syncwarp();
    const int f = lane >> 3, j = lane & 7;
    const double *leaf = s + 128 * f + 8 * ((f + 1) >> 1) + j;
    double r = leaf[0];
#pragma unroll
    for (int i = 8; i < 128; i += 8) r += leaf[i];
    r = r + __shfl_xor_sync(0xffffffffu, r, 1);
    r = r + __shfl_xor_sync(0xffffffffu, r, 2);
    r = r + __shfl_xor_sync(0xffffffffu, r, 4);
    r = r + __shfl_xor_sync(0xffffffffu, r, 8);
    const double tot = r + __shfl_xor_sync(0xffffffffu, r, 16);
#pragma unroll
    for (int k = 0; k < M / 32; k++) {
        const double post = x[k] / tot;
        dst[32 * k + lane] = scale_by_power ? post * scale : post;
    }
}

// Section-wise softmax denoiser of amp() for an arbitrary (dense) design matrix (sparc_ldpc.py:214-219):
// beta = sqrt(n P_l) softmax_section(s sqrt(n P_l) / tau^2), and sum(beta^2) per section.  One warp per section;
// codewords with active[b] == 0 are left untouched (per-codeword early stop).
__global__ void section_softmax_kernel(const double *__restrict__ s, const double *__restrict__ Pl,
                                       const double *__restrict__ tau2, const unsigned char *__restrict__ active, int L,
                                       int M, int n, double *__restrict__ beta, double *__restrict__ sumsq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int b = blockIdx.y, l = blockIdx.x * wpb + warp;
    if (l >= L || (active && !active[b])) return;
    const double rt = sqrt((double)n * Pl[l]), c2 = rt / tau2[b];
    const double *src = s + ((size_t)b * L + l) * M;
    double *dst = beta + ((size_t)b * L + l) * M;
    double m = -INFINITY;
    for (int j = lane; j < M; j += 32) m = fmax(m, src[j] * c2);
    for (int d = 16; d; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    double sum = 0.0;
    for (int j = lane; j < M; j += 32) sum += exp(src[j] * c2 - m);
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const double sc = rt / sum;
    double sq = 0.0;
    for (int j = lane; j < M; j += 32) {
        const double v = exp(src[j] * c2 - m) * sc;
        dst[j] = v;
        sq += v * v;
    }
    for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    if (lane == 0) sumsq[(size_t)b * L + l] = sq;  // per section: the caller adds them up deterministically
}

// idx = argmax of each section, first maximum wins (np.argmax).  One warp per section.
__global__ void argmax_kernel(const double *__restrict__ beta, long beta_stride, int count, int M, int *idx,
                              long idx_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * wpb + warp;
    if (i >= count) return;
    const double *src = beta + (size_t)b * beta_stride + (size_t)i * M;
    double best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < M; j += 32) {
        const double v = src[j];
        if (v > best || (bi == 0x7fffffff)) { best = v; bi = j; }
    }
    for (int d = 16; d; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) idx[(size_t)b * idx_stride + i] = bi;
}

__global__ void llr2idx_kernel(const double *__restrict__ llr, long llr_stride, int count, int logM, int *idx,
                               long idx_stride) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double *src = llr + (size_t)b * llr_stride + (size_t)i * logM;
    int v = 0;
    for (int j = 0; j < logM; j++) v = (v << 1) | (src[j] < 0.0 ? 1 : 0);  // MSB first (:331-339), bit = LLR<0 (:352)
    idx[(size_t)b * idx_stride + i] = v;
}

__global__ void count_errors_kernel(const int *__restrict__ a, const int *__restrict__ t, int count, int *errs) {
    __shared__ int red[32];
    const int b = blockIdx.x;
    int e = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) e += __popc(a[(size_t)b * count + i] ^ t[(size_t)b * count + i]);
    for (int d = 16; d; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < (blockDim.x + 31) / 32; i++) s += red[i];
        errs[b] = s;
    }
}

// amp_exit.py:85-106: a protected section with exactly one entry above the threshold is hard decided.
__global__ void peel_kernel(const double *__restrict__ post, int L, int M, int ls, double thr, int *hard_idx, int *act,
                            int *nact) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *hs = reinterpret_cast<int *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int b = blockIdx.x;
    for (int l = warp; l < L; l += nw) {
        int cnt = 0, pos = -1;
        if (l >= L - ls) {
            const double *src = post + ((size_t)b * L + l) * M;
            for (int j = lane; j < M; j += 32)
                if (src[j] > thr) { cnt++; pos = j; }
            for (int d = 16; d; d >>= 1) {
                cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
                pos = max(pos, __shfl_xor_sync(0xffffffffu, pos, d));
            }
        }
        if (lane == 0) hs[l] = (cnt == 1) ? pos : -1;
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) hard_idx[(size_t)b * L + l] = hs[l];
    if (threadIdx.x == 0) {
        int na = 0;
        for (int l = 0; l < L; l++)
            if (hs[l] < 0) act[(size_t)b * L + na++] = l;
        nact[b] = na;
        for (int l = na; l < L; l++) act[(size_t)b * L + l] = 0;
    }
}

// beta[b][l*M + j] = (j == idx[b][l]) ? sqrt(n*Pl[l]) : 0   (sparc_ldpc.py:840-843; idx < 0 leaves the section zero)
__global__ void onehot_beta_kernel(const int *__restrict__ idx, const double *__restrict__ Pl, int n, int L, int M,
                                   double *__restrict__ beta) {
    const int b = blockIdx.y;
    const size_t LM = (size_t)L * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < LM; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i / M), j = (int)(i % M);
        beta[(size_t)b * LM + i] = (idx[(size_t)b * L + l] == j) ? sqrt((double)n * Pl[l]) : 0.0;
    }
}

// Systematic QC-LDPC encoder (ldpc/py/ldpc.py:790-850), one CTA per codeword, bits as bytes in shared memory.
// proto[Mp][Np] (shift or -1); info[B][Kp*z] -> x[B][Np*z]; toff = the single odd-multiplicity shift of column Kp.
__global__ void ldpc_encode_kernel(const int *__restrict__ proto, int Mp, int Np, int z, int toff,
                                   const unsigned char *__restrict__ info, unsigned char *__restrict__ x) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Kp = Np - Mp, b = blockIdx.x;
    unsigned char *xs = smem_raw;              // [Np][z]
    unsigned char *syn = xs + (size_t)Np * z;  // [Mp][z] systematic part of every block-row
    for (int i = threadIdx.x; i < Kp * z; i += blockDim.x) xs[i] = info[(size_t)b * Kp * z + i] & 1;
    __syncthreads();
    for (int i = threadIdx.x; i < Mp * z; i += blockDim.x) {
        const int r = i / z, k = i % z;
        unsigned char acc = 0;
        for (int c = 0; c < Kp; c++) {
            const int sft = proto[r * Np + c];
            if (sft >= 0) acc ^= xs[c * z + (k + sft) % z];  // np.roll(x_c, -shift)[k]
        }
        syn[i] = acc;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < z; k += blockDim.x) {  // first parity block: roll(sum of all rows, toff)
        const int src = ((k - toff) % z + z) % z;
        unsigned char acc = 0;
        for (int r = 0; r < Mp; r++) acc ^= syn[r * z + src];
        xs[Kp * z + k] = acc;
    }
    __syncthreads();
    for (int r = 0; r < Mp - 1; r++) {  // back-substitution down the dual diagonal
        for (int k = threadIdx.x; k < z; k += blockDim.x) {
            unsigned char acc = syn[r * z + k];
            for (int c = 0; c <= r; c++) {
                const int sft = proto[r * Np + Kp + c];
                if (sft >= 0) acc ^= xs[(Kp + c) * z + (k + sft) % z];
            }
            xs[(Kp + r + 1) * z + k] = acc;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < Np * z; i += blockDim.x) x[(size_t)b * Np * z + i] = xs[i];
}

// idx[b][i] = MSB-first value of logM consecutive bits (bytes 0/1)   (sparc_ldpc.py:317-341)
__global__ void bits2idx_kernel(const unsigned char *__restrict__ bits, long bits_stride, int count, int logM, int *idx,
                                long idx_stride) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned char *src = bits + (size_t)b * bits_stride + (size_t)i * logM;
    int v = 0;
    for (int j = 0; j < logM; j++) v = (v << 1) | (src[j] & 1);
    idx[(size_t)b * idx_stride + i] = v;
}

// np.histogram(E[X == +-1], bins=edges): left-closed bins, last bin closed on the right.
__global__ void exit_hist_kernel(const double *__restrict__ E, const int *__restrict__ X, int len,
                                 const double *__restrict__ edges, int nedges, unsigned long long *counts) {
    const int b = blockIdx.y, nb = nedges - 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const double x = E[(size_t)b * len + i];
    const int sgn = X[(size_t)b * len + i];
    if (!(x >= edges[0]) || !(x <= edges[nb])) return;
    int lo = 0, hi = nb;  // largest j with edges[j] <= x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= x) lo = mid; else hi = mid;
    }
    if (x == edges[nb]) lo = nb - 1;
    const int which = (sgn == 1) ? 0 : ((sgn == -1) ? 1 : -1);
    if (which < 0) return;
    atomicAdd(counts + ((size_t)b * 2 + which) * nb + lo, 1ull);
}

}  // namespace sb

using namespace sb;

extern "C" int sb_sp2bp_llr_batch(const double *beta, long beta_stride, int beta_first, const int *sections,
                                  const int *nsec, int L_stride, int first_sec, int out_first, int count, int M, int n,
                                  const double *Pl, int B, double *p, double *llr, long out_stride, void *stream) {
    if (!beta || !Pl || !llr || B < 0 || M < 2 || (M & (M - 1)) || M > 1024)
        return fail(SB_EINVAL, "sb_sp2bp_llr_batch: bad argument%s", "");
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_sp2bp_llr_batch: sections and nsec go together%s", "");
    const int maxcount = sections ? L_stride : count;
    if (B == 0 || maxcount <= 0) return SB_OK;
    if (M >= 64 && !knob("SB_HANDOFF_V1")) {  // 16 sections per CTA, one lane per (section, bit) chain
        const int logM = ilog2(M);
        int R = knob("SB_SP2BP_R") ? atoi(knob("SB_SP2BP_R")) : 8;
        if (R != 4 && R != 8 && R != 16) R = 8;  // only these are instantiated  // 8 measured best (1.26 ms vs 1.54 at 16, 1.38 at 4; 2.03 one warp per section)
        dim3 grid16((maxcount + R - 1) / R, B);
        const size_t smem16 = sizeof(double) * R * (size_t)(M + 1);
        if (R == 4) {
            if (smem16 > 48 * 1024)
                SB_CUDA(cudaFuncSetAttribute(sp2bp_llr_kernel16<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            sp2bp_llr_kernel16<4><<<grid16, 64, smem16, (cudaStream_t)stream>>>(beta, beta_stride, beta_first, sections, nsec,
                                                                             L_stride, first_sec, out_first, count, M,
                                                                             logM, n, Pl, p, llr, out_stride);
        } else if (R == 8) {
            if (smem16 > 48 * 1024)
                SB_CUDA(cudaFuncSetAttribute(sp2bp_llr_kernel16<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            sp2bp_llr_kernel16<8><<<grid16, 128, smem16, (cudaStream_t)stream>>>(beta, beta_stride, beta_first, sections, nsec,
                                                                              L_stride, first_sec, out_first, count, M,
                                                                              logM, n, Pl, p, llr, out_stride);
        } else {
            if (smem16 > 48 * 1024)
                SB_CUDA(cudaFuncSetAttribute(sp2bp_llr_kernel16<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            sp2bp_llr_kernel16<16><<<grid16, 256, smem16, (cudaStream_t)stream>>>(beta, beta_stride, beta_first, sections, nsec,
                                                                               L_stride, first_sec, out_first, count, M,
                                                                               logM, n, Pl, p, llr, out_stride);
        }
        SB_LAUNCHED();
        return SB_OK;
    }
    const int wpb = 8;
    dim3 grid((maxcount + wpb - 1) / wpb, B);
    const size_t smem = sizeof(double) * (size_t)wpb * M;
    if (smem > 48 * 1024)
        SB_CUDA(cudaFuncSetAttribute(sp2bp_llr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sp2bp_llr_kernel<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(beta, beta_stride, beta_first, sections, nsec,
                                                                    L_stride, first_sec, out_first, count, M,
                                                                    ilog2(M), n, Pl, p, llr, out_stride);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_bp2sp_prior_batch(const double *app, int ls, const double *beta_prev, int L, int M, int n,
                                    const double *Pl, int mode, int B, double *beta_init, void *stream) {
    const int scale_by_power = mode & SB_PRIOR_SCALE, input_is_prob = (mode & SB_PRIOR_FROM_PROB) ? 1 : 0;
    if (!beta_init || !Pl || B < 0 || ls < 0 || ls > L || M < 2 || (M & (M - 1)) || M > 1024 || (ls > 0 && !app) ||
        (ls < L && !beta_prev))
        return fail(SB_EINVAL, "sb_bp2sp_prior_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    const int wpb = 8;
    dim3 grid((L + wpb - 1) / wpb, B);
    const size_t smem = sizeof(double) * (size_t)wpb * (M + 16);
    if (smem > 48 * 1024)
        SB_CUDA(cudaFuncSetAttribute(bp2sp_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (M == 512 && !knob("SB_HANDOFF_V1")) {
        SB_CUDA(cudaFuncSetAttribute(bp2sp_prior_kernel512, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bp2sp_prior_kernel512<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(app, ls, beta_prev, L, n, Pl, scale_by_power,
                                                                             input_is_prob, beta_init);
    } else
        bp2sp_prior_kernel<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(app, ls, beta_prev, L, M, ilog2(M), n, Pl,
                                                                          scale_by_power, input_is_prob, beta_init);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_section_softmax_batch(const double *s, const double *Pl, const double *tau2,
                                        const unsigned char *active, int L, int M, int n, int B, double *beta,
                                        double *sumsq, void *stream) {
    if (!s || !Pl || !tau2 || !beta || !sumsq || B < 0 || L <= 0 || M <= 0)
        return fail(SB_EINVAL, "sb_section_softmax_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    SB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * (size_t)B * L, (cudaStream_t)stream));
    const int wpb = 8;
    dim3 grid((L + wpb - 1) / wpb, B);
    section_softmax_kernel<<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(s, Pl, tau2, active, L, M, n, beta, sumsq);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_argmax_batch(const double *beta, long beta_stride, int count, int M, int B, int *idx,
                               long idx_stride, void *stream) {
    if (!beta || !idx || B < 0 || count < 0 || M < 1) return fail(SB_EINVAL, "sb_argmax_batch: bad argument%s", "");
    if (B == 0 || count == 0) return SB_OK;
    const int wpb = 8;
    dim3 grid((count + wpb - 1) / wpb, B);
    argmax_kernel<<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(beta, beta_stride, count, M, idx, idx_stride);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_llr2idx_batch(const double *llr, long llr_stride, int count, int M, int B, int *idx,
                                long idx_stride, void *stream) {
    if (!llr || !idx || B < 0 || count < 0 || M < 2 || (M & (M - 1)))
        return fail(SB_EINVAL, "sb_llr2idx_batch: bad argument%s", "");
    if (B == 0 || count == 0) return SB_OK;
    dim3 grid((count + 127) / 128, B);
    llr2idx_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(llr, llr_stride, count, ilog2(M), idx, idx_stride);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_count_errors_batch(const int *a, const int *t, int count, int B, int *errs, void *stream) {
    if (!a || !t || !errs || B < 0 || count < 0) return fail(SB_EINVAL, "sb_count_errors_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    count_errors_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(a, t, count, errs);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_threshold_peel_batch(const double *post, int L, int M, int ls, double threshold, int B,
                                       int *hard_idx, int *act, int *nact, void *stream) {
    if (!post || !hard_idx || !act || !nact || B < 0 || L <= 0 || ls < 0 || ls > L)
        return fail(SB_EINVAL, "sb_threshold_peel_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    const size_t smem = sizeof(int) * (size_t)L;
    if (smem > 48 * 1024)
        SB_CUDA(cudaFuncSetAttribute(peel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    peel_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(post, L, M, ls, threshold, hard_idx, act, nact);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_onehot_beta_batch(const int *idx, const double *Pl, int n, int L, int M, int B, double *beta,
                                    void *stream) {
    if (!idx || !Pl || !beta || B < 0 || L <= 0 || M <= 0) return fail(SB_EINVAL, "sb_onehot_beta_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    const size_t LM = (size_t)L * M;
    dim3 grid((unsigned)((LM + 255) / 256 > 1184 ? 1184 : (LM + 255) / 256), B);
    onehot_beta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, Pl, n, L, M, beta);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_ldpc_encode_batch(const int *proto, int Mp, int Np, int z, int toff, const unsigned char *info, int B,
                                    unsigned char *x, void *stream) {
    if (!proto || !info || !x || Mp <= 0 || Np <= Mp || z <= 0 || B < 0)
        return fail(SB_EINVAL, "sb_ldpc_encode_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    const size_t smem = (size_t)(Np + Mp) * z;
    if (smem > 227 * 1024) return fail(SB_EINVAL, "sb_ldpc_encode_batch: code too large for shared memory%s", "");
    if (smem > 48 * 1024)
        SB_CUDA(cudaFuncSetAttribute(ldpc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ldpc_encode_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(proto, Mp, Np, z, toff, info, x);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_bits2idx_batch(const unsigned char *bits, long bits_stride, int count, int M, int B, int *idx,
                                 long idx_stride, void *stream) {
    if (!bits || !idx || B < 0 || count < 0 || M < 2 || (M & (M - 1)))
        return fail(SB_EINVAL, "sb_bits2idx_batch: bad argument%s", "");
    if (B == 0 || count == 0) return SB_OK;
    dim3 grid((count + 127) / 128, B);
    bits2idx_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(bits, bits_stride, count, ilog2(M), idx, idx_stride);
    SB_LAUNCHED();
    return SB_OK;
}

extern "C" int sb_exit_hist_batch(const double *E, const int *X, int len, const double *edges, int nedges, int B,
                                  long long *counts, void *stream) {
    if (!E || !X || !edges || !counts || len < 0 || nedges < 2 || B < 0)
        return fail(SB_EINVAL, "sb_exit_hist_batch: bad argument%s", "");
    if (B == 0 || len == 0) return SB_OK;
    SB_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * (size_t)B * 2 * (nedges - 1), (cudaStream_t)stream));
    dim3 grid((len + 255) / 256, B);
    exit_hist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(E, X, len, edges, nedges,
                                                            reinterpret_cast<unsigned long long *>(counts));
    SB_LAUNCHED();
    return SB_OK;
}
