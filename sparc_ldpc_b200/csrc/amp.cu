// amp.cu -- host side of the SPARC design operator / AMP entry points: table construction and dispatch.
// The kernels live in amp_impl.cuh and are instantiated per section size in amp_inst_*.cu.
#include "amp_impl.cuh"

namespace sb {

// out[b][k] = y[b][k] + sign * (sum_l c_l * sgn(l,k) * H_M[lo(l,k), idx_l]) / sqrt(n): the transform of a one-hot
// section is +-c exactly, so this equals the reference's Ab(beta_onehot) bit for bit (sections ascending).
__global__ void onehot_kernel(const uint16_t *__restrict__ fwd, int L, int n, int SBQ, const int *__restrict__ idx,
                              const double *__restrict__ Pl, const double *__restrict__ y, double sign,
                              double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *sidx = reinterpret_cast<int *>(smem_raw);
    double *coef = reinterpret_cast<double *>(sidx + ((L + 1) & ~1));
    const int b = blockIdx.y;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        sidx[l] = idx[(size_t)b * L + l];
        coef[l] = sqrt((double)n * Pl[l]);
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double rt_n = sqrt((double)n);
    const uint32_t lomask = (1u << SBQ) - 1u;
    double acc = 0.0;
    for (int l = 0; l < L; l++) {
        const int j = sidx[l];
        if (j < 0) continue;
        const uint32_t e = __ldg(fwd + (size_t)l * n + k);
        const int neg = ((e >> SBQ) & 1) ^ (__popc(((e & lomask) >> 2) & (uint32_t)j) & 1);
        acc += neg ? -coef[l] : coef[l];
    }
    const double x = acc / rt_n;
    const double base = y ? y[(size_t)b * n + k] : 0.0;
    out[(size_t)b * n + k] = (sign < 0) ? (base - x) : (base + x);
}

// one butterfly stage of the in-place Walsh-Hadamard transform (ldpc/sparc_ldpc.py:19-29): stride h
__global__ void fht_stage_kernel(double *x, long N, long h) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N / 2) return;
    const long j = (t / h) * 2 * h + (t % h);
    const double a = x[j], b = x[j + h];
    x[j] = a + b;
    x[j + h] = a - b;
}

extern template int launch_amp<1>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<2>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<3>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<4>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<5>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<6>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<7>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<8>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<9>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);
extern template int launch_amp<10>(const sb_operator *, AmpArgs, int, int, const double *, double *, cudaStream_t);

static int dispatch(const sb_operator *op, AmpArgs a, int B, int which, const double *in, double *out,
                    cudaStream_t st) {
    switch (op->logM) {
#define SB_CASE(l) \
    case l:        \
        return launch_amp<l>(op, a, B, which, in, out, st);
        SB_CASE(1) SB_CASE(2) SB_CASE(3) SB_CASE(4) SB_CASE(5) SB_CASE(6) SB_CASE(7) SB_CASE(8) SB_CASE(9) SB_CASE(10)
#undef SB_CASE
    }
    return fail(SB_EINVAL, "unsupported section size M = 2^%s%ld", "", op->logM);
}

static int sign_bit(int logM) { return logM + 2 + (logM >= 10 ? 2 : 3); }

}  // namespace sb

using namespace sb;

extern "C" int sb_operator_create(const uint32_t *ordering, int L, int M, int n, sb_operator **out) {
    if (!ordering || !out || L <= 0 || n <= 0 || M < 2 || (M & (M - 1)) || M > 1024 || n >= 65534)
        return fail(SB_EINVAL, "sb_operator_create: bad shape%s (M=%ld)", "", M);
    int w = 1;
    while (w < (M + 1 > n + 1 ? M + 1 : n + 1)) w <<= 1;  // sparc_ldpc.py:54,110
    sb_operator *op = new sb_operator();
    op->L = L; op->M = M; op->n = n; op->logM = ilog2(M); op->w = w;
    op->H = w / M;
    op->Hp = op->H < 16 ? 16 : op->H;
    op->NB = op->Hp / 16;
    if (op->NB > 128) { delete op; return fail(SB_EINVAL, "sb_operator_create: w/M too large%s (%ld)", "", op->H); }
    const int logH = ilog2(op->H), SBQ = sign_bit(op->logM);
    op->pre = ((size_t)n * 4 <= 65535) ? 1 : 0;  // inverse-table entries as int32 byte offsets when they fit in 16 bits
    op->G8 = (L + 7) / 8;
    const size_t nf = (size_t)L * n, ni = (size_t)L * M * op->Hp, n8 = (size_t)op->G8 * n * 8;
    uint16_t *hf = (uint16_t *)malloc(nf * 2), *hi = (uint16_t *)malloc(ni * 2), *h8 = (uint16_t *)calloc(n8, 2);
    if (!hf || !hi || !h8) { free(hf); free(hi); free(h8); delete op; return fail(SB_ENOMEM, "sb_operator_create: host alloc%s", ""); }
    const uint16_t empty = (uint16_t)(op->pre ? n * 4 : n);  // the zero word zs[n]
    for (size_t i = 0; i < ni; i++) hi[i] = empty;
    for (int l = 0; l < L; l++)
        for (int k = 0; k < n; k++) {
            const uint32_t r = ordering[(size_t)l * n + k];
            if (r == 0 || r >= (uint32_t)w) { free(hf); free(hi); free(h8); delete op; return fail(SB_EINVAL, "ordering entry out of [1,w)%s", ""); }
            const uint32_t lo = r % M, hiw = r / M;
            uint32_t c = 0;  // visit position = bit reversal of the block index over log2(H) bits
            for (int bbit = 0; bbit < logH; bbit++) c |= ((hiw >> bbit) & 1u) << (logH - 1 - bbit);
            const uint16_t fe = (uint16_t)((lo << 2) | ((__builtin_popcount(hiw) & 1) << SBQ));
            hf[(size_t)l * n + k] = fe;
            h8[((size_t)(l >> 3) * n + k) * 8 + (l & 7)] = fe;
            hi[((size_t)l * M + lo) * op->Hp + c] = (uint16_t)(op->pre ? k * 4 : k);
        }
    op->fwd = nullptr; op->inv = nullptr; op->fwd8 = nullptr;
    cudaError_t e1 = cudaMalloc(&op->fwd, nf * 2), e2 = cudaMalloc(&op->inv, ni * 2), e3 = cudaMalloc(&op->fwd8, n8 * 2);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(op->fwd); cudaFree(op->inv); cudaFree(op->fwd8); free(hf); free(hi); free(h8); delete op;
        return fail(SB_ENOMEM, "sb_operator_create: cudaMalloc failed%s", "");
    }
    e1 = cudaMemcpy(op->fwd, hf, nf * 2, cudaMemcpyHostToDevice);
    e2 = cudaMemcpy(op->inv, hi, ni * 2, cudaMemcpyHostToDevice);
    e3 = cudaMemcpy(op->fwd8, h8, n8 * 2, cudaMemcpyHostToDevice);
    free(hf); free(hi); free(h8);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { sb_operator_destroy(op); return fail(SB_ECUDA, "sb_operator_create: copy failed%s", ""); }
    *out = op;
    return SB_OK;
}

extern "C" void sb_operator_destroy(sb_operator *op) {
    if (!op) return;
    cudaFree(op->fwd);
    cudaFree(op->fwd8);
    cudaFree(op->inv);
    delete op;
}

static AmpArgs base_args(const sb_operator *op, const int *sections, const int *nsec) {
    AmpArgs a;
    memset(&a, 0, sizeof(a));
    a.fwd = op->fwd; a.fwd8 = op->fwd8; a.inv = op->inv; a.sections = sections; a.nsec = nsec;
    a.L = op->L; a.n = op->n; a.Hp = op->Hp; a.NB = op->NB;
    return a;
}

extern "C" int sb_amp_batch(const sb_operator *op, const double *y, const double *Pl, const double *beta0,
                            const int *sections, const int *nsec, int B, int T, int mode, double *beta, int *iters,
                            int *n_exec, unsigned *flags, double *tau2_trace, double *scratch, void *stream) {
    if (!op || !y || !Pl || !beta || !iters || !n_exec || !flags || B < 0 || T < 0)
        return fail(SB_EINVAL, "sb_amp_batch: null argument%s", "");
    if (mode != SB_AMP_STRICT && mode != SB_AMP_FAST) return fail(SB_EINVAL, "sb_amp_batch: unknown mode%s %ld", "", mode);
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_amp_batch: sections and nsec go together%s", "");
    if (mode == SB_AMP_FAST && op->pre && !scratch) return fail(SB_EINVAL, "sb_amp_batch: FAST mode needs a [B][n] scratch%s", "");
    if (B == 0) return SB_OK;
    AmpArgs a = base_args(op, sections, nsec);
    a.zscratch = scratch;
    a.y = y; a.Pl = Pl; a.beta0 = beta0; a.beta = beta; a.tau2_trace = tau2_trace;
    a.iters = iters; a.n_exec = n_exec; a.flags = flags; a.T = T;
    return dispatch(op, a, B, mode == SB_AMP_FAST ? 3 : 0, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int sb_Ab_batch(const sb_operator *op, const double *beta, const int *sections, const int *nsec, int B,
                           double *out, void *stream) {
    if (!op || !beta || !out || B < 0) return fail(SB_EINVAL, "sb_Ab_batch: null argument%s", "");
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_Ab_batch: sections and nsec go together%s", "");
    if (B == 0) return SB_OK;
    AmpArgs a = base_args(op, sections, nsec);
    return dispatch(op, a, B, 1, beta, out, (cudaStream_t)stream);
}

extern "C" int sb_Az_batch(const sb_operator *op, const double *z, const int *sections, const int *nsec, int B,
                           double *out, void *stream) {
    if (!op || !z || !out || B < 0) return fail(SB_EINVAL, "sb_Az_batch: null argument%s", "");
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_Az_batch: sections and nsec go together%s", "");
    if (B == 0) return SB_OK;
    AmpArgs a = base_args(op, sections, nsec);
    return dispatch(op, a, B, 2, z, out, (cudaStream_t)stream);
}

// pyfht.fht_inplace replacement (sparc_ldpc.py:14-29): host pointer, length a power of two, transformed in place
// with the stages in the reference's order (strides N/2 ... 1), hence bit-identical to it.
extern "C" int sb_fht_inplace_host(double *x, long N) {
    if (!x || N <= 0 || (N & (N - 1))) return fail(SB_EINVAL, "sb_fht_inplace_host: length must be a power of two%s (%ld)", "", N);
    if (N == 1) return SB_OK;
    double *d = nullptr;
    SB_CUDA(cudaMalloc(&d, sizeof(double) * N));
    cudaError_t e = cudaMemcpy(d, x, sizeof(double) * N, cudaMemcpyHostToDevice);
    for (long h = N >> 1; h && e == cudaSuccess; h >>= 1) {
        fht_stage_kernel<<<(unsigned)((N / 2 + 255) / 256), 256>>>(d, N, h);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(x, d, sizeof(double) * N, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SB_ECUDA, "sb_fht_inplace_host: %s", cudaGetErrorString(e));
    return SB_OK;
}

extern "C" int sb_onehot_apply_batch(const sb_operator *op, const int *idx, const double *Pl, const double *y,
                                     double sign, int B, double *out, void *stream) {
    if (!op || !idx || !Pl || !out || B < 0) return fail(SB_EINVAL, "sb_onehot_apply_batch: null argument%s", "");
    if (B == 0) return SB_OK;
    const int nt = 256;
    dim3 grid((op->n + nt - 1) / nt, B);
    const size_t smem = sizeof(int) * ((op->L + 1) & ~1) + sizeof(double) * op->L;
    if (smem > 48 * 1024)
        SB_CUDA(cudaFuncSetAttribute(onehot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    onehot_kernel<<<grid, nt, smem, (cudaStream_t)stream>>>(op->fwd, op->L, op->n, sign_bit(op->logM), idx, Pl, y, sign,
                                                          out);
    SB_LAUNCHED();
    return SB_OK;
}
