// bp.cu -- batched flooding belief propagation over a Tanner graph (sm_100a, fp64).
//
// Replaces ldpc/src/c_ldpc.c of the reference: sumprod2 (:138-206, the default decoder), sumprod (:32-113),
// minsum (:339-381), Lxor (:234-251), Lxfb (:294-314).  One CTA decodes one codeword; all Nmsg messages stay
// in shared memory for the whole decode.  The arithmetic of every node follows the reference literally
// (same operand order, exp then log without log1p, signbit/fmin/fabs) so that results differ from the CPU
// only through the last-ulp differences between CUDA's and glibc's exp/log/tanh/atanh.
//
// Message layout: the reference stores messages check-major (check c owns cdeg[c] consecutive slots).  A
// thread-per-check sweep over that layout hits the same bank 8 times at dc = 20, so internally slot
// (check c, port k) lives at  classoff[dc] + k * classsize[dc] + rank_in_class(c):  consecutive checks of
// one degree touch consecutive doubles, and for quasi-cyclic codes consecutive variables do too.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace sb {

__host__ __device__ __forceinline__ double lxor(double L1, double L2, int corr) {
    double L = (signbit(L1) == signbit(L2)) ? 1.0 : -1.0;  // c_ldpc.c:239-242
    L *= fmin(fabs(L1), fabs(L2));                          // :243
    if (corr) {
        L += log(1 + exp(-fabs(L1 + L2)));  // :246
        L -= log(1 + exp(-fabs(L1 - L2)));  // :247
    }
    return L;
}

// FAST rule (SB_BP_SUMPROD2_FAST): the same node update with the two correction terms log(1 + exp(-|x|)) of every
// Lxor evaluated in single precision (ex2 / lg2 special-function units) and added to the fp64 sign-min term.  The
// terms are bounded by ln 2, so the absolute error per Lxor is ~1e-7; messages, variable-node sums, app and the
// stop test stay fp64.  ~8 fp64-pipe instructions per Lxor instead of ~90 (2 exp + 2 log in fp64).
__device__ __forceinline__ double lxor_fast(double L1, double L2) {
    double L = (signbit(L1) == signbit(L2)) ? 1.0 : -1.0;
    L *= fmin(fabs(L1), fabs(L2));
    const float s = (float)fabs(L1 + L2), d = (float)fabs(L1 - L2);
    const float c = __logf(1.0f + __expf(-s)) - __logf(1.0f + __expf(-d));
    return L + (double)c;
}

template <int RULE>
__device__ __forceinline__ double lxor_rule(double L1, double L2) {
    if (RULE == SB_BP_SUMPROD2_FAST) return lxor_fast(L1, L2);
    return lxor(L1, L2, RULE == SB_BP_SUMPROD2);
}

struct BpArgs {
    const int *voff, *vpos, *cbase, *cstride, *cdeg;
    const double *ch;
    double *app;
    int *it;
    int Nv, Nc, Nmsg, max_it;
    double factor;
};

// Forward/backward extrinsics of one check in place (c_ldpc.c:294-314); sb = this thread's scratch column.
__device__ __forceinline__ double check_fb(double *msg, int base, int stride, int dc, double *sb, int sbs, int corr) {
    double bk = msg[base + (dc - 1) * stride];
    sb[(dc - 1) * sbs] = bk;
    for (int k = dc - 2; k >= 0; k--) {
        bk = lxor(bk, msg[base + k * stride], corr);  // b[k] = Lxor(b[k+1], L[k])      (:305)
        sb[k * sbs] = bk;
    }
    double f = msg[base];
    msg[base] = sb[sbs];  // L[0] = b[1]                                                   (:309)
    for (int k = 1; k < dc - 1; k++) {
        const double Lk = msg[base + k * stride];
        msg[base + k * stride] = lxor(f, sb[(k + 1) * sbs], corr);  // L[k] = Lxor(f[k-1], b[k+1])   (:311)
        f = lxor(f, Lk, corr);                                      // f[k] = Lxor(f[k-1], L[k])      (:303)
    }
    msg[base + (dc - 1) * stride] = f;  // L[dc-1] = f[dc-2]
    return bk;                          // b[0]
}

// The same node update with the backward values b[k] in a per-thread LOCAL array (L1-resident: 768 threads x
// DCMAX doubles fit the L1 left beside the messages) instead of a shared-memory scratch column: one thread per
// check fits (768 threads at the headline code instead of 384 with two checks each), and twice as many warps
// hide the exp / log latency.  Operands and their order are those of check_fb, i.e. of Lxfb
// (c_ldpc.c:294-314): results are bit-identical.
template <int RULE, int DCMAX>
__device__ __forceinline__ double check_fb_reg(double *msg, int base, int stride, int dc) {
    double bs[DCMAX];
    double bk = msg[base + (dc - 1) * stride];
    bs[dc - 1] = bk;
#pragma unroll 1
    for (int k = dc - 2; k >= 0; k--) {
        bk = lxor_rule<RULE>(bk, msg[base + k * stride]);  // b[k] = Lxor(b[k+1], L[k])      (:305)
        bs[k] = bk;
    }
    double f = msg[base];
    msg[base] = bs[1];  // L[0] = b[1]                                                   (:309)
#pragma unroll 1
    for (int k = 1; k < dc - 1; k++) {
        const double Lk = msg[base + k * stride];
        msg[base + k * stride] = lxor_rule<RULE>(f, bs[k + 1]);  // L[k] = Lxor(f[k-1], b[k+1])   (:311)
        f = lxor_rule<RULE>(f, Lk);                              // f[k] = Lxor(f[k-1], L[k])      (:303)
    }
    msg[base + (dc - 1) * stride] = f;  // L[dc-1] = f[dc-2]
    return bk;                          // b[0]
}

template <int RULE, int DCMAX>
__global__ void __launch_bounds__(768, 1) bp_kernel_reg(BpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *msg = reinterpret_cast<double *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const double *ch = a.ch + (size_t)b * a.Nv;
    double *app = a.app + (size_t)b * a.Nv;
    for (int i = tid; i < a.Nmsg; i += NT) msg[i] = 0.0;  // calloc, c_ldpc.c:164
    __syncthreads();
    int it;
    for (it = 0; it < a.max_it; it++) {
        for (int v = tid; v < a.Nv; v += NT) {  // variable nodes (c_ldpc.c:171-178)
            const int p0 = a.voff[v], p1 = a.voff[v + 1];
            double aggr = ch[v];
            for (int p = p0; p < p1; p++) aggr += msg[a.vpos[p]];
            for (int p = p0; p < p1; p++) {
                const int s = a.vpos[p];
                msg[s] = aggr - msg[s];
            }
            app[v] = aggr;
        }
        __syncthreads();
        int unsat = 0;
        for (int c = tid; c < a.Nc; c += NT) {
            const int dc = a.cdeg[c], base = a.cbase[c], stride = a.cstride[c];
            const double tot = check_fb_reg<RULE, DCMAX>(msg, base, stride, dc);
            if (tot <= 0.0) unsat = 1;  // c_ldpc.c:191
            if (RULE == SB_BP_MINSUM)
                for (int k = 0; k < dc; k++) msg[base + k * stride] *= a.factor;  // :370-371
        }
        if (!__syncthreads_or(unsat)) break;  // c_ldpc.c:196
    }
    if (tid == 0) a.it[b] = it;
}

template <int RULE>
__global__ void bp_kernel(BpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *msg = reinterpret_cast<double *>(smem_raw);
    double *scr = msg + a.Nmsg;
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const double *ch = a.ch + (size_t)b * a.Nv;
    double *app = a.app + (size_t)b * a.Nv;
    for (int i = tid; i < a.Nmsg; i += NT) msg[i] = 0.0;  // calloc, c_ldpc.c:164
    __syncthreads();
    int it;
    for (it = 0; it < a.max_it; it++) {
        // variable nodes (c_ldpc.c:171-178): aggr = ch + sum of inputs; out = aggr - in
        for (int v = tid; v < a.Nv; v += NT) {
            const int p0 = a.voff[v], p1 = a.voff[v + 1];
            double aggr = ch[v];
            for (int p = p0; p < p1; p++) aggr += msg[a.vpos[p]];
            for (int p = p0; p < p1; p++) {
                const int s = a.vpos[p];
                msg[s] = aggr - msg[s];
            }
            app[v] = aggr;
        }
        __syncthreads();
        int unsat = 0;
        for (int c = tid; c < a.Nc; c += NT) {
            const int dc = a.cdeg[c], base = a.cbase[c], stride = a.cstride[c];
            if (RULE == SB_BP_SUMPROD) {  // c_ldpc.c:76-102
                double aggr = 1.0;
                for (int k = 0; k < dc; k++) {
                    const double t = tanh(msg[base + k * stride] / 2.0);
                    msg[base + k * stride] = t;
                    aggr *= t;
                }
                if (2.0 * atanh(aggr) <= 0.0) unsat = 1;
                for (int k = 0; k < dc; k++) msg[base + k * stride] = 2.0 * atanh(aggr / msg[base + k * stride]);
            } else {
                const double tot = check_fb(msg, base, stride, dc, scr + tid, NT, RULE == SB_BP_SUMPROD2);
                if (tot <= 0.0) unsat = 1;  // c_ldpc.c:191
                if (RULE == SB_BP_MINSUM)
                    for (int k = 0; k < dc; k++) msg[base + k * stride] *= a.factor;  // :370-371
            }
        }
        if (!__syncthreads_or(unsat)) break;  // c_ldpc.c:196
    }
    if (tid == 0) a.it[b] = it;
}

__global__ void lxor_kernel(double L1, double L2, int corr, double *out) { *out = lxor(L1, L2, corr); }

__global__ void lxfb_kernel(double *L, int dc, int corr, double *scratch, double *tot) {
    *tot = check_fb(L, 0, 1, dc, scratch, 1, corr);
}

static int bp_threads(const sb_graph *g, size_t *smem_out) {
    const size_t budget = 227 * 1024, msgb = sizeof(double) * (size_t)g->Nmsg;
    if (msgb + sizeof(double) * g->dcmax * 32 > budget) return -1;
    int ntmax = (int)((budget - msgb) / (sizeof(double) * g->dcmax));
    if (ntmax > 1024) ntmax = 1024;
    ntmax = (ntmax / 32) * 32;
    const int passes = (g->Nc + ntmax - 1) / ntmax;
    int nt = ((g->Nc + passes - 1) / passes + 31) / 32 * 32;
    if (nt > ntmax) nt = ntmax;
    if (nt < 32) nt = 32;
    *smem_out = msgb + sizeof(double) * (size_t)g->dcmax * nt;
    return nt;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_graph_create(const long *vdeg, const long *cdeg, const long *intrlv, int Nv, int Nc, int Nmsg,
                               sb_graph **out) {
    if (!vdeg || !cdeg || !intrlv || !out || Nv <= 0 || Nc <= 0 || Nmsg <= 0)
        return fail(SB_EINVAL, "sb_graph_create: bad argument%s", "");
    std::vector<int> voff(Nv + 1), vpos(Nmsg), cbase(Nc), cstride(Nc), cd(Nc), e2i(Nmsg);
    long sv = 0, sc = 0;
    int dcmax = 0, dvmax = 0;
    for (int v = 0; v < Nv; v++) {
        if (vdeg[v] < 0) return fail(SB_EINVAL, "sb_graph_create: negative degree%s", "");
        voff[v] = (int)sv; sv += vdeg[v];
        if (vdeg[v] > dvmax) dvmax = (int)vdeg[v];
    }
    voff[Nv] = (int)sv;
    for (int c = 0; c < Nc; c++) {
        if (cdeg[c] < 2) return fail(SB_EINVAL, "sb_graph_create: check degree < 2%s", "");
        sc += cdeg[c];
        if (cdeg[c] > dcmax) dcmax = (int)cdeg[c];
    }
    if (sv != Nmsg || sc != Nmsg) return fail(SB_EINVAL, "sb_graph_create: degree sums differ from Nmsg%s (%ld)", "", Nmsg);
    std::vector<int> cnt(dcmax + 1, 0), off(dcmax + 2, 0), rank(Nc);
    for (int c = 0; c < Nc; c++) rank[c] = cnt[cdeg[c]]++;
    for (int d = 0; d <= dcmax; d++) off[d + 1] = off[d] + d * cnt[d];
    long m = 0;
    for (int c = 0; c < Nc; c++) {
        const int d = (int)cdeg[c];
        cd[c] = d; cbase[c] = off[d] + rank[c]; cstride[c] = cnt[d];
        for (int k = 0; k < d; k++) e2i[m++] = cbase[c] + k * cstride[c];
    }
    for (int p = 0; p < Nmsg; p++) {
        if (intrlv[p] < 0 || intrlv[p] >= Nmsg) return fail(SB_EINVAL, "sb_graph_create: interleaver entry out of range%s", "");
        vpos[p] = e2i[intrlv[p]];
    }
    sb_graph *g = new sb_graph();
    memset(g, 0, sizeof(*g));
    g->Nv = Nv; g->Nc = Nc; g->Nmsg = Nmsg; g->dcmax = dcmax; g->dvmax = dvmax;
    struct { int **dst; std::vector<int> *src; } items[] = {{&g->voff, &voff}, {&g->vpos, &vpos}, {&g->cbase, &cbase},
                                                            {&g->cstride, &cstride}, {&g->cdeg, &cd}, {&g->ext2int, &e2i}};
    for (auto &itx : items) {
        const size_t bytes = sizeof(int) * itx.src->size();
        if (cudaMalloc(itx.dst, bytes) != cudaSuccess ||
            cudaMemcpy(*itx.dst, itx.src->data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            sb_graph_destroy(g);
            return fail(SB_ENOMEM, "sb_graph_create: device allocation/copy failed%s", "");
        }
    }
    *out = g;
    return SB_OK;
}

extern "C" void sb_graph_destroy(sb_graph *g) {
    if (!g) return;
    cudaFree(g->voff); cudaFree(g->vpos); cudaFree(g->cbase); cudaFree(g->cstride); cudaFree(g->cdeg); cudaFree(g->ext2int);
    delete g;
}

extern "C" int sb_bp_batch(const sb_graph *g, int rule, const double *ch, int B, double *app, int *it, int max_it,
                           double minsum_factor, void *stream) {
    if (!g || !ch || !app || !it || B < 0 || max_it < 0) return fail(SB_EINVAL, "sb_bp_batch: bad argument%s", "");
    if (B == 0) return SB_OK;
    size_t smem = 0;
    const int nt = bp_threads(g, &smem);
    if (nt < 0) return fail(SB_EINVAL, "sb_bp_batch: graph does not fit in shared memory%s (Nmsg=%ld)", "", g->Nmsg);
    BpArgs a{g->voff, g->vpos, g->cbase, g->cstride, g->cdeg, ch, app, it, g->Nv, g->Nc, g->Nmsg, max_it, minsum_factor};
    cudaStream_t st = (cudaStream_t)stream;
    // local-array variant (no shared-memory scratch): one thread per check, up to 768; SB_BP_SCRATCH=1 forces the scratch variant
    const size_t msgb = sizeof(double) * (size_t)g->Nmsg;
    const bool reg_ok = rule != SB_BP_SUMPROD && g->dcmax <= 24 && msgb <= 227 * 1024 && !knob("SB_BP_SCRATCH");
    if (rule == SB_BP_SUMPROD2_FAST && !reg_ok) rule = SB_BP_SUMPROD2;  // the scratch-column kernel has no FAST variant
    if (reg_ok) {
        int ntr = ((g->Nc + 31) / 32) * 32;
        if (ntr > 768) {  // several checks per thread: balanced passes (768 threads = 85 registers each)
            const int passes = (g->Nc + 767) / 768;
            ntr = (((g->Nc + passes - 1) / passes) + 31) / 32 * 32;
        }
#define SB_REG(r, d)                                                                                              \
    do {                                                                                                          \
        SB_CUDA(cudaFuncSetAttribute(bp_kernel_reg<r, d>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msgb)); \
        bp_kernel_reg<r, d><<<B, ntr, msgb, st>>>(a);                                                            \
    } while (0)
#define SB_REG_DC(r)                                    \
    do {                                                \
        if (g->dcmax <= 8) SB_REG(r, 8);                \
        else if (g->dcmax <= 16) SB_REG(r, 16);         \
        else SB_REG(r, 24);                             \
    } while (0)
        if (rule == SB_BP_SUMPROD2) SB_REG_DC(SB_BP_SUMPROD2);
        else if (rule == SB_BP_SUMPROD2_FAST) SB_REG_DC(SB_BP_SUMPROD2_FAST);
        else if (rule == SB_BP_MINSUM) SB_REG_DC(SB_BP_MINSUM);
        else return fail(SB_EINVAL, "sb_bp_batch: unknown rule%s %ld", "", rule);
#undef SB_REG_DC
#undef SB_REG
        SB_LAUNCHED();
        return SB_OK;
    }
    switch (rule) {
#define SB_RULE(r)                                                                                            \
    case r:                                                                                                   \
        SB_CUDA(cudaFuncSetAttribute(bp_kernel<r>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        bp_kernel<r><<<B, nt, smem, st>>>(a);                                                                \
        break;
        SB_RULE(SB_BP_SUMPROD2) SB_RULE(SB_BP_SUMPROD) SB_RULE(SB_BP_MINSUM)
#undef SB_RULE
        default:
            return fail(SB_EINVAL, "sb_bp_batch: unknown rule%s %ld", "", rule);
    }
    SB_LAUNCHED();
    return SB_OK;
}

// ------------------------------------------------------------------ reference FFI (host pointers, one codeword)
namespace {
std::mutex g_mu;
struct Cached {
    sb_graph *g = nullptr;
    uint64_t key = 0;
    double *d_ch = nullptr, *d_app = nullptr;
    int *d_it = nullptr;
    int cap = 0;
} g_cache;

uint64_t hash_arrays(const long *a, int na, const long *b, int nb, const long *c, int nc) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const long *p, int n) {
        for (int i = 0; i < n; i++) { h ^= (uint64_t)p[i] + 0x9e3779b97f4a7c15ull; h *= 1099511628211ull; }
        h ^= (uint64_t)n; h *= 1099511628211ull;
    };
    mix(a, na); mix(b, nb); mix(c, nc);
    return h;
}

int ref_decode(int rule, double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app,
               double factor) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!ch || !vdeg || !cdeg || !intrlv || !app) return -1;
    const uint64_t key = hash_arrays(vdeg, Nv, cdeg, Nc, intrlv, Nmsg);
    if (!g_cache.g || g_cache.key != key) {
        sb_graph_destroy(g_cache.g);
        g_cache.g = nullptr;
        if (sb_graph_create(vdeg, cdeg, intrlv, Nv, Nc, Nmsg, &g_cache.g) != SB_OK) return -1;
        g_cache.key = key;
    }
    if (g_cache.cap < Nv) {
        cudaFree(g_cache.d_ch); cudaFree(g_cache.d_app); cudaFree(g_cache.d_it);
        g_cache.cap = 0;
        if (cudaMalloc(&g_cache.d_ch, sizeof(double) * Nv) != cudaSuccess ||
            cudaMalloc(&g_cache.d_app, sizeof(double) * Nv) != cudaSuccess ||
            cudaMalloc(&g_cache.d_it, sizeof(int)) != cudaSuccess)
            return -1;
        g_cache.cap = Nv;
    }
    int it = -1;
    if (cudaMemcpy(g_cache.d_ch, ch, sizeof(double) * Nv, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (sb_bp_batch(g_cache.g, rule, g_cache.d_ch, 1, g_cache.d_app, g_cache.d_it, SB_MAX_ITCOUNT, factor, nullptr) != SB_OK)
        return -1;
    if (cudaMemcpy(app, g_cache.d_app, sizeof(double) * Nv, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(&it, g_cache.d_it, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return it;
}
}  // namespace

extern "C" int sumprod(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app) {
    return ref_decode(SB_BP_SUMPROD, ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, 1.0);
}
extern "C" int sumprod2(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app) {
    return ref_decode(SB_BP_SUMPROD2, ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, 1.0);
}
extern "C" int minsum(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app,
                      double correction_factor) {
    return ref_decode(SB_BP_MINSUM, ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, correction_factor);
}

// The two scalar helpers run on the device as well, so that code.Lxor / code.Lxfb (ldpc.py:932-943) return
// exactly what the batched decoder computes.  They fail loudly (NaN) when no CUDA device is usable.
extern "C" double Lxor(double L1, double L2, int corr_flag) {
    std::lock_guard<std::mutex> lk(g_mu);
    double *d = nullptr, r = NAN;
    if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) { fail(SB_ECUDA, "Lxor: no CUDA device%s", ""); return NAN; }
    lxor_kernel<<<1, 1>>>(L1, L2, corr_flag, d);
    g_launches.fetch_add(1);
    if (cudaMemcpy(&r, d, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) r = NAN;
    cudaFree(d);
    return r;
}

extern "C" double Lxfb(double *L, long dc, int corr_flag) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!L || dc < 2 || dc > 4096) { fail(SB_EINVAL, "Lxfb: bad degree%s %ld", "", dc); return NAN; }
    double *d = nullptr, r = NAN;
    if (cudaMalloc(&d, sizeof(double) * (2 * dc + 1)) != cudaSuccess) { fail(SB_ECUDA, "Lxfb: no CUDA device%s", ""); return NAN; }
    if (cudaMemcpy(d, L, sizeof(double) * dc, cudaMemcpyHostToDevice) == cudaSuccess) {
        lxfb_kernel<<<1, 1>>>(d, (int)dc, corr_flag, d + dc, d + 2 * dc);
        g_launches.fetch_add(1);
        if (cudaMemcpy(L, d, sizeof(double) * dc, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(&r, d + 2 * dc, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
            r = NAN;
    }
    cudaFree(d);
    return r;
}

// ---- measured instruction ceiling of the check-node function (bench.py roofline_bp) ---------------------------
// Every thread runs 4 independent chains of `iters` dependent Lxor evaluations on register operands: no memory
// traffic, all SMs, 8 CTAs of 256 threads per SM.  Lxor/s of this loop is what the fp64 / special-function pipes
// can sustain for the rule -- the ceiling the resident BP kernel is compared with.
namespace sb {
template <int RULE>
__global__ void __launch_bounds__(256) lxor_peak_kernel(int iters, double seed, double *out) {
    double a0 = seed + 1e-3 * threadIdx.x, a1 = -a0 * 0.7, a2 = a0 * 1.3, a3 = -a0 * 0.4;
    const double b0 = 0.9 + 1e-4 * blockIdx.x, b1 = -1.7, b2 = 2.3, b3 = -0.35;
    for (int i = 0; i < iters; i++) {
        a0 = lxor_rule<RULE>(a0, b0) + b1;
        a1 = lxor_rule<RULE>(a1, b1) + b2;
        a2 = lxor_rule<RULE>(a2, b2) + b3;
        a3 = lxor_rule<RULE>(a3, b3) + b0;
    }
    if (a0 + a1 + a2 + a3 == 12345.678) out[0] = a0;  // keeps the chains alive
}
}  // namespace sb

extern "C" int sb_bp_lxor_peak(int rule, double *lxor_per_s) {
    if (!lxor_per_s || (rule != SB_BP_SUMPROD2 && rule != SB_BP_SUMPROD2_FAST && rule != SB_BP_MINSUM))
        return sb::fail(SB_EINVAL, "sb_bp_lxor_peak: rule must be SB_BP_SUMPROD2, SB_BP_SUMPROD2_FAST or SB_BP_MINSUM%s", "");
    int dev = 0, nsm = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    double *out = nullptr;
    SB_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0));
    SB_CUDA(cudaEventCreate(&e1));
    const int grid = nsm * 8, iters = rule == SB_BP_SUMPROD2 ? 512 : 4096;
    float ms = 0.f;
    for (int rep = 0; rep < 2; rep++) {  // first pass warms up
        SB_CUDA(cudaEventRecord(e0));
        if (rule == SB_BP_SUMPROD2) sb::lxor_peak_kernel<SB_BP_SUMPROD2><<<grid, 256>>>(iters, 0.37, out);
        else if (rule == SB_BP_SUMPROD2_FAST) sb::lxor_peak_kernel<SB_BP_SUMPROD2_FAST><<<grid, 256>>>(iters, 0.37, out);
        else sb::lxor_peak_kernel<SB_BP_MINSUM><<<grid, 256>>>(iters, 0.37, out);
        SB_LAUNCHED();
        SB_CUDA(cudaEventRecord(e1));
        SB_CUDA(cudaEventSynchronize(e1));
        SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *lxor_per_s = 4.0 * iters * 256.0 * grid / (ms * 1e-3);
    return SB_OK;
}
