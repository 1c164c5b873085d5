// amp.cu -- SPARC design operator and AMP decoder kernels (sm_100a, fp64).
//
// Replaces ldpc/sparc_ldpc.py:32-147 (sub_fht / block_sub_fht / sparc_transforms[_shorter]) and
// :189-222 (amp) of the reference.  The reference zero-pads every section to w = 2^ceil(log2(n+1))
// and runs a w-point Walsh-Hadamard transform; because only the last M columns of H_w are used,
//     (A beta)[k]  = (1/sqrt n) sum_l sgn(l,k) * FHT_M(beta_l)[lo(l,k)]
//     (A^T z)_l    = (1/sqrt n) FHT_M(fold_l(z)),
// with r = ordering[l][k], lo = r mod M, hi = r div M, sgn = (-1)^popcount(hi) and fold_l the signed
// butterfly tree over the w/M blocks.  Both identities are exact in floating point when the adds are
// done in the reference's order, which these kernels do: the fold tree is evaluated in the order of the
// large-stride butterflies, the M-point stages run from stride M/2 down to 1, and sections are
// accumulated into A beta in ascending order (sparc_ldpc.py:123-126).
//
// Kernel structure: ONE persistent CTA per codeword runs the whole AMP loop (all T iterations) in a
// single launch.  z and the A beta accumulator live in shared memory; beta streams through HBM once in
// and once out per iteration (the section-wise softmax needs only the section itself); the lookup tables
// are shared by the whole batch and stay L2-resident:
//     inv  u16 [L][M][Hp]   for bin lo of section l: the k (as a byte offset into z) of every block, in the
//                           visit order of the fold tree; empty blocks point at a zero word
//     fwd  u16 [L][n]       for row k of section l: lo*8 | sgn << SB   (section lists)
//     fwd8 u16 [L/8][n][8]  the same entries with 8 sections interleaved: one 16-byte load per (group, k)
// Per group of W sections each team (<= one warp) transforms one section and leaves FHT_M(beta_l) in shared
// memory TWICE, as +F and as -F at byte distance 2^SB, so that the signed gather of A beta is a single
// LDS at offset (entry) followed by a DADD -- no sign handling in the inner loop.
#include "common.cuh"

namespace sb {

__host__ __device__ __forceinline__ int zpad(int n) { return (n + 2) & ~1; }  // n values + zero word, even

template <int LOGM>
struct TeamCfg {
    static constexpr int M = 1 << LOGM;
    static constexpr int TEAM = (M >= 128) ? 32 : (M >= 4 ? M / 4 : 1);  // lanes cooperating on one section
    static constexpr int EPT = M / TEAM;                                  // elements per lane
    static constexpr int SPR = (LOGM >= 10) ? 4 : 8;                      // section slots per F region
    static constexpr int SB = LOGM + 3 + ((LOGM >= 10) ? 2 : 3);          // sign bit of a fwd entry (<= 15)
    static constexpr int REGION = 2 << SB;                                // bytes: +F half then -F half
};

// shuffle mask of the calling lane's team (teams of one warp may diverge from each other)
template <int TEAM>
__device__ __forceinline__ unsigned team_mask() {
    if constexpr (TEAM >= 32) {
        return 0xffffffffu;
    } else {
        return ((1u << TEAM) - 1u) << (((threadIdx.x & 31) / TEAM) * TEAM);
    }
}

// M-point Walsh-Hadamard transform of one section held by a team: element j = e*TEAM + q lives in
// x[e] of team lane q.  Stage order = strides M/2 ... 1 (ldpc/sparc_ldpc.py:19-29): (a, b) -> (a+b, a-b).
template <int LOGM>
__device__ __forceinline__ void fht_team(double (&x)[TeamCfg<LOGM>::EPT], int q, unsigned tmask) {
    constexpr int TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
#pragma unroll
    for (int s = EPT / 2; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            if ((i & s) == 0) {
                double a = x[i], b = x[i + s];
                x[i] = a + b;
                x[i + s] = a - b;
            }
        }
    }
#pragma unroll
    for (int d = TEAM / 2; d >= 1; d >>= 1) {
        // lane with bit d set holds x[ij]: new = partner - mine = partner + (-mine); negation = sign-bit xor
        const int sgn = (q & d) ? (int)0x80000000 : 0;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const double p = __shfl_xor_sync(tmask, x[e], d);
            const double mine = __hiloint2double(__double2hiint(x[e]) ^ sgn, __double2loint(x[e]));
            x[e] = p + mine;
        }
    }
}

// z value addressed by an inverse-table entry.  PRE: entries are byte offsets (k*8) instead of indices.
template <bool PRE>
__device__ __forceinline__ double zs_at(const double *zs, uint32_t k) {
    if (PRE) return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(zs) + k);
    return zs[k];
}

// One block of 16 inverse-table entries in visit order; every tree node is (left - right), i.e.
// v <- v[:half] - v[half:] of the reference's large-stride butterflies.  Empty slots point at zs[n] = 0.
template <bool PRE>
__device__ __forceinline__ double fold16(const uint4 p0, const uint4 p1, const double *zs) {
    const uint32_t wds[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    double v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[2 * i] = zs_at<PRE>(zs, wds[i] & 0xFFFFu);
        v[2 * i + 1] = zs_at<PRE>(zs, wds[i] >> 16);
    }
#pragma unroll
    for (int s = 1; s < 16; s <<= 1) {
#pragma unroll
        for (int i = 0; i < 16; i += 2 * s) v[i] = v[i] - v[i + s];
    }
    return v[0];
}

// One bin with NB > 1 blocks: binary-counter merge of the block subtrees (left - right at every level).
template <bool PRE>
__device__ __forceinline__ double fold_bin_multi(const uint16_t *__restrict__ tab, int NB, const double *zs) {
    double st[8];
    double val = 0.0;
    for (int c = 0; c < NB; c++) {
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(tab + c * 16));
        const uint4 p1 = __ldg(reinterpret_cast<const uint4 *>(tab + c * 16 + 8));
        val = fold16<PRE>(p0, p1, zs);
        int cc = c, lvl = 0;
#pragma unroll
        for (int l = 0; l < 7; l++) {
            if (cc & 1) {
                val = st[l] - val;
                cc >>= 1;
                lvl = l + 1;
            } else {
                break;
            }
        }
#pragma unroll
        for (int l = 0; l < 8; l++)
            if (l == lvl) st[l] = val;
    }
    return val;
}

// fold_l(z) for the EPT bins of this lane.  NB == 1 (w/M <= 16, the headline shapes): the two 16-byte table
// loads of bin e+1 are issued before bin e is reduced.
template <int LOGM, bool PRE>
__device__ __forceinline__ void fold_section(double (&x)[TeamCfg<LOGM>::EPT], const uint16_t *__restrict__ tab,
                                             int Hp, int NB, int q, const double *zs) {
    constexpr int TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
    if (NB == 1) {
        const uint4 *t4 = reinterpret_cast<const uint4 *>(tab);  // Hp == 16: two uint4 per bin
        uint4 c0 = __ldg(t4 + 2 * q), c1 = __ldg(t4 + 2 * q + 1);
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            uint4 n0 = c0, n1 = c1;
            if (e + 1 < EPT) {
                n0 = __ldg(t4 + 2 * ((e + 1) * TEAM + q));
                n1 = __ldg(t4 + 2 * ((e + 1) * TEAM + q) + 1);
            }
            x[e] = fold16<PRE>(c0, c1, zs);
            c0 = n0;
            c1 = n1;
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) x[e] = fold_bin_multi<PRE>(tab + (size_t)(e * TEAM + q) * Hp, NB, zs);
    }
}

struct AmpArgs {
    const uint16_t *fwd, *fwd8, *inv;
    const double *y, *Pl, *beta0;
    const int *sections, *nsec;
    double *beta, *tau2_trace;
    int *iters, *n_exec;
    unsigned *flags;
    int L, n, Hp, NB, T;
};

// byte offset of section slot `slot` (+F copy) inside the F area
template <int LOGM>
__device__ __forceinline__ int slot_offset(int slot) {
    using C = TeamCfg<LOGM>;
    return (slot / C::SPR) * C::REGION + (slot % C::SPR) * (C::M * 8);
}

// mode 0: AMP iteration (fold -> FHT -> softmax -> store beta -> FHT -> +-F)
// mode 1: operator only (load beta -> FHT -> +-F)            [prologue z = y - A beta0, sb_Ab_batch]
template <int LOGM, bool PRE>
__device__ __forceinline__ void section_phase(int mode, bool first_zero, const AmpArgs &a, const double *bsrc,
                                              double *bdst, int sec, int q, const double *zs, char *Fbytes, int slot,
                                              double inv_rt_n, double rt_npl, double tau2, double &sq, double &gmax,
                                              double &lmin) {
    constexpr int M = TeamCfg<LOGM>::M, TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
    double x[EPT];
    const unsigned tmask = team_mask<TEAM>();
    if (mode == 0) {
        fold_section<LOGM, PRE>(x, a.inv + ((size_t)sec * M) * a.Hp, a.Hp, a.NB, q, zs);
        fht_team<LOGM>(x, q, tmask);
        const double c2 = rt_npl / tau2;
        double m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const double b = first_zero ? 0.0 : bsrc[e * TEAM + q];
            const double s = b + x[e] * inv_rt_n;  // s = beta + A^T z          (sparc_ldpc.py:213)
            x[e] = s * c2;                         // u = s sqrt(n P_l)/tau^2    (:215)
            m = fmax(m, x[e]);
        }
#pragma unroll
        for (int d = TEAM / 2; d >= 1; d >>= 1) m = fmax(m, __shfl_xor_sync(tmask, m, d));
        gmax = fmax(gmax, m);
        lmin = fmin(lmin, m);
        double sum = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            x[e] = exp(x[e] - m);  // section max instead of the reference's global max (:216): same softmax
            sum += x[e];
        }
#pragma unroll
        for (int d = TEAM / 2; d >= 1; d >>= 1) sum += __shfl_xor_sync(tmask, sum, d);
        const double sc = rt_npl / sum;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            x[e] *= sc;  // beta = sqrt(n P_l) softmax(u)    (:218-219)
            sq += x[e] * x[e];
            bdst[e * TEAM + q] = x[e];
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            x[e] = bsrc[e * TEAM + q];
            if (bdst != nullptr) bdst[e * TEAM + q] = x[e];
        }
    }
    fht_team<LOGM>(x, q, tmask);
    double *Fp = reinterpret_cast<double *>(Fbytes + slot_offset<LOGM>(slot));
    double *Fn = reinterpret_cast<double *>(Fbytes + slot_offset<LOGM>(slot) + (TeamCfg<LOGM>::REGION >> 1));
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        Fp[e * TEAM + q] = x[e];
        Fn[e * TEAM + q] = -x[e];
    }
}

// acc[k] += sum over the group's sections (ascending) of sgn * F[lo]     (sparc_ldpc.py:123-126, :70)
// generic section lists: one u16 row per section
template <int LOGM>
__device__ __forceinline__ void gather_phase(const uint16_t *__restrict__ fwd, int n, int nvalid, const int *sec_s,
                                             const char *Fbytes, double *acc_s) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double acc = acc_s[k];
        int tm = 0;
        for (; tm + 4 <= nvalid; tm += 4) {
            uint32_t e[4];
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = __ldg(fwd + (size_t)sec_s[tm + i] * n + k);
#pragma unroll
            for (int i = 0; i < 4; i++)
                acc += *reinterpret_cast<const double *>(Fbytes + slot_offset<LOGM>(tm + i) + e[i]);
        }
        for (; tm < nvalid; tm++)
            acc += *reinterpret_cast<const double *>(Fbytes + slot_offset<LOGM>(tm) +
                                                     __ldg(fwd + (size_t)sec_s[tm] * n + k));
        acc_s[k] = acc;
    }
}

// all sections in order and groups aligned to 8: the 8 entries of (group, k) are one 16-byte load from the
// interleaved table.  KB rows per thread are in flight together and chunk c+1 is loaded while chunk c is used.
template <int LOGM>
__device__ __forceinline__ void gather_phase8(const uint16_t *__restrict__ fwd8, int n, int g0, int nvalid,
                                              const char *Fbytes, double *acc_s) {
    constexpr int KB = 4;
    const uint4 *tab = reinterpret_cast<const uint4 *>(fwd8) + (size_t)(g0 >> 3) * n;
    const int nch = (nvalid + 7) >> 3, NT = blockDim.x;
    for (int k0 = threadIdx.x; k0 < n; k0 += KB * NT) {
        double acc[KB];
        uint4 w[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) {
            const int k = k0 + j * NT;
            acc[j] = (k < n) ? acc_s[k] : 0.0;
            w[j] = (k < n) ? __ldg(tab + k) : make_uint4(0, 0, 0, 0);
        }
        for (int c = 0; c < nch; c++) {
            uint4 wn[KB];
#pragma unroll
            for (int j = 0; j < KB; j++) {
                const int k = k0 + j * NT;
                wn[j] = (c + 1 < nch && k < n) ? __ldg(tab + (size_t)(c + 1) * n + k) : make_uint4(0, 0, 0, 0);
            }
            const char *F = Fbytes + slot_offset<LOGM>(c * 8);  // c*8 is a multiple of SPR: region start
            const bool full = (c * 8 + 8 <= nvalid);
#pragma unroll
            for (int j = 0; j < KB; j++) {
                const uint32_t wds[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (full || c * 8 + i < nvalid) {
                        const uint32_t e = (i & 1) ? (wds[i >> 1] >> 16) : (wds[i >> 1] & 0xFFFFu);
                        acc[j] += *reinterpret_cast<const double *>(F + slot_offset<LOGM>(i) + e);
                    }
                }
                w[j] = wn[j];
            }
        }
#pragma unroll
        for (int j = 0; j < KB; j++) {
            const int k = k0 + j * NT;
            if (k < n) acc_s[k] = acc[j];
        }
    }
}

// One pass over all active sections: section_phase per team, then the gather per group.
template <int LOGM, bool PRE>
__device__ __forceinline__ void operator_pass(int mode, bool first_zero, const AmpArgs &a, const double *bsrc,
                                              double *bdst, const int *act, int La, double *zs, double *acc_s,
                                              char *Fbytes, int *sec_s, int W, double inv_rt_n, double nd,
                                              double tau2, double &sq, double &gmax, double &lmin) {
    constexpr int M = TeamCfg<LOGM>::M, TEAM = TeamCfg<LOGM>::TEAM;
    const int tm = threadIdx.x / TEAM, q = threadIdx.x % TEAM;
    for (int k = threadIdx.x; k < a.n; k += blockDim.x) acc_s[k] = 0.0;
    for (int g0 = 0; g0 < La; g0 += W) {
        const int sidx = g0 + tm;
        const bool valid = (tm < W) && (sidx < La);
        const int sec = valid ? (act ? act[sidx] : sidx) : 0;
        if (q == 0 && tm < W) sec_s[tm] = sec;
        if (valid) {
            const double rt_npl = sqrt(nd * a.Pl[sec]);
            section_phase<LOGM, PRE>(mode, first_zero, a, bsrc ? bsrc + (size_t)sidx * M : nullptr,
                                     bdst ? bdst + (size_t)sidx * M : nullptr, sec, q, zs, Fbytes, tm, inv_rt_n,
                                     rt_npl, tau2, sq, gmax, lmin);
        }
        __syncthreads();
        const int nvalid = min(W, La - g0);
        if (act == nullptr && (W & 7) == 0)
            gather_phase8<LOGM>(a.fwd8, a.n, g0, nvalid, Fbytes, acc_s);
        else
            gather_phase<LOGM>(a.fwd, a.n, nvalid, sec_s, Fbytes, acc_s);
        __syncthreads();
    }
}

// shared-memory carve-up: zs[zpad(n)] | acc[n (+1 to keep 16-byte alignment)] | F area | red[40] | sec[W]
template <int LOGM>
struct Smem {
    double *zs, *acc, *red;
    char *F;
    int *sec;
    __device__ Smem(unsigned char *raw, int n, int W) {
        zs = reinterpret_cast<double *>(raw);
        acc = zs + zpad(n);
        F = reinterpret_cast<char *>(acc + ((n + 1) & ~1));
        const int regions = (W + TeamCfg<LOGM>::SPR - 1) / TeamCfg<LOGM>::SPR;
        red = reinterpret_cast<double *>(F + (size_t)regions * TeamCfg<LOGM>::REGION);
        sec = reinterpret_cast<int *>(red + 40);
    }
};

template <int LOGM>
static size_t amp_smem(int n, int W) {
    const int regions = (W + TeamCfg<LOGM>::SPR - 1) / TeamCfg<LOGM>::SPR;
    return sizeof(double) * ((size_t)zpad(n) + ((n + 1) & ~1) + 40) + (size_t)regions * TeamCfg<LOGM>::REGION +
           sizeof(int) * (size_t)(W + 2);
}

template <int LOGM, bool PRE>
__global__ void __launch_bounds__(512, 1) amp_kernel(AmpArgs a, int W) {
    constexpr int M = TeamCfg<LOGM>::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, b = blockIdx.x;
    Smem<LOGM> sm(smem_raw, n, W);
    double *zs = sm.zs, *acc_s = sm.acc, *red = sm.red;
    if (threadIdx.x == 0) zs[n] = 0.0;

    const int La = a.nsec ? a.nsec[b] : a.L;
    const int *act = a.sections ? a.sections + (size_t)b * a.L : nullptr;
    const double *y = a.y + (size_t)b * n;
    double *beta = a.beta + (size_t)b * a.L * M;
    const double nd = (double)n;
    const double rt_n = sqrt(nd), inv_rt_n = 1.0 / rt_n;
    double sq = 0.0, gmax = -INFINITY, lmin = INFINITY;

    if (La <= 0) {  // the reference never calls amp() on an empty section set (sparc_ldpc.py:1015)
        if (threadIdx.x == 0) {
            a.iters[b] = 0;
            a.n_exec[b] = 0;
            a.flags[b] = 0;
        }
        return;
    }

    // P = sum of the active sections' power (np.sum(Pl), sparc_ldpc.py:190)
    double pl = 0.0;
    for (int i = threadIdx.x; i < La; i += blockDim.x) pl += a.Pl[act ? act[i] : i];
    const double P = block_sum(pl, red);

    if (a.beta0 != nullptr) {  // z = y - A beta0   (sparc_ldpc.py:197-198)
        operator_pass<LOGM, PRE>(1, false, a, a.beta0 + (size_t)b * a.L * M, beta, act, La, zs, acc_s, sm.F, sm.sec, W,
                                 inv_rt_n, nd, 1.0, sq, gmax, lmin);
        for (int k = threadIdx.x; k < n; k += blockDim.x) zs[k] = y[k] - acc_s[k] / rt_n;
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) zs[k] = y[k];
    }
    __syncthreads();

    bool first_zero = (a.beta0 == nullptr);
    double last_tau = 0.0;
    unsigned flags = 0;
    int t = 0, executed = 0;
    for (t = 0; t < a.T; t++) {
        double part = 0.0;
        for (int k = threadIdx.x; k < n; k += blockDim.x) part += zs[k] * zs[k];
        const double tau = sqrt(block_sum(part, red) / nd);  // (:203)
        if (tau == last_tau) {                               // exact-equality stop (:204)
            flags |= SB_AMP_STOPPED;
            break;
        }
        last_tau = tau;
        const double tau2 = tau * tau;
        if (a.tau2_trace != nullptr && threadIdx.x == 0) a.tau2_trace[(size_t)b * a.T + t] = tau2;
        sq = 0.0;
        gmax = -INFINITY;
        lmin = INFINITY;
        operator_pass<LOGM, PRE>(0, first_zero, a, beta, beta, act, La, zs, acc_s, sm.F, sm.sec, W, inv_rt_n, nd, tau2,
                                 sq, gmax, lmin);
        first_zero = false;
        const double sumsq = block_sum(sq, red);
        const double gm = block_max(gmax, red);
        const double lm = -block_max(-lmin, red);
        if (gm - lm > 745.13) flags |= SB_AMP_REF_NAN;
        const double ons = P - sumsq / nd;  // (:220)
        for (int k = threadIdx.x; k < n; k += blockDim.x) zs[k] = (y[k] - acc_s[k] / rt_n) + (zs[k] / tau2) * ons;
        __syncthreads();
        executed++;
    }
    if (first_zero) {  // T == 0 or stop before the first update: beta is the zero vector
        for (int i = threadIdx.x; i < La * M; i += blockDim.x) beta[i] = 0.0;
    }
    if (threadIdx.x == 0) {
        a.iters[b] = (t < a.T) ? t : (a.T > 0 ? a.T - 1 : 0);
        a.n_exec[b] = executed;
        a.flags[b] = flags;
    }
}

// A_S beta for a batch (sparc_ldpc.py:143-144): out = acc / sqrt(n)
template <int LOGM, bool PRE>
__global__ void __launch_bounds__(512, 1) Ab_kernel(AmpArgs a, int W, const double *beta_in, double *out) {
    constexpr int M = TeamCfg<LOGM>::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, b = blockIdx.x;
    Smem<LOGM> sm(smem_raw, n, W);
    const int La = a.nsec ? a.nsec[b] : a.L;
    const int *act = a.sections ? a.sections + (size_t)b * a.L : nullptr;
    double sq = 0, gmax = 0, lmin = 0;
    const double nd = (double)n, rt_n = sqrt(nd);
    operator_pass<LOGM, PRE>(1, false, a, beta_in + (size_t)b * a.L * M, nullptr, act, La, sm.zs, sm.acc, sm.F, sm.sec,
                             W, 1.0 / rt_n, nd, 1.0, sq, gmax, lmin);
    for (int k = threadIdx.x; k < n; k += blockDim.x) out[(size_t)b * n + k] = sm.acc[k] / rt_n;
}

// A_S^T z for a batch (sparc_ldpc.py:145-146): one team per section
template <int LOGM, bool PRE>
__global__ void __launch_bounds__(512, 1) Az_kernel(AmpArgs a, int W, const double *z_in, double *out) {
    constexpr int M = TeamCfg<LOGM>::M, TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, b = blockIdx.x;
    const int NTM = blockDim.x / TEAM;
    double *zs = reinterpret_cast<double *>(smem_raw);
    const int La = a.nsec ? a.nsec[b] : a.L;
    const int *act = a.sections ? a.sections + (size_t)b * a.L : nullptr;
    for (int k = threadIdx.x; k < n; k += blockDim.x) zs[k] = z_in[(size_t)b * n + k];
    if (threadIdx.x == 0) zs[n] = 0.0;
    __syncthreads();
    const double rt_n = sqrt((double)n);
    const int tm = threadIdx.x / TEAM, q = threadIdx.x % TEAM;
    for (int sidx = tm; sidx < La; sidx += NTM) {  // whole teams leave together: shuffles stay converged
        const int sec = act ? act[sidx] : sidx;
        double x[EPT];
        fold_section<LOGM, PRE>(x, a.inv + ((size_t)sec * M) * a.Hp, a.Hp, a.NB, q, zs);
        fht_team<LOGM>(x, q, team_mask<TEAM>());
#pragma unroll
        for (int e = 0; e < EPT; e++) out[(size_t)b * a.L * M + (size_t)sidx * M + e * TEAM + q] = x[e] / rt_n;
    }
}

// out[b][k] = y[b][k] + sign * (sum_l c_l * sgn(l,k) * H_M[lo(l,k), idx_l]) / sqrt(n): the transform of a one-hot
// section is +-c exactly, so this equals the reference's Ab(beta_onehot) bit for bit (sections ascending).
__global__ void onehot_kernel(const uint16_t *__restrict__ fwd, int L, int n, int SB, const int *__restrict__ idx,
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
    const uint32_t lomask = (1u << SB) - 1u;
    double acc = 0.0;
    for (int l = 0; l < L; l++) {
        const int j = sidx[l];
        if (j < 0) continue;
        const uint32_t e = __ldg(fwd + (size_t)l * n + k);
        const int neg = ((e >> SB) & 1) ^ (__popc(((e & lomask) >> 3) & (uint32_t)j) & 1);
        acc += neg ? -coef[l] : coef[l];
    }
    const double x = acc / rt_n;
    const double base = y ? y[(size_t)b * n + k] : 0.0;
    out[(size_t)b * n + k] = (sign < 0) ? (base - x) : (base + x);
}

// threads per CTA: as many teams as fit in shared memory (<= 512 threads), never more teams than sections
template <int LOGM>
static int pick_threads(int n, int L, size_t *smem_out, int *W_out) {
    constexpr int TEAM = TeamCfg<LOGM>::TEAM;
    int nt = 512;
    const char *env = getenv("SB_AMP_THREADS");
    if (env) nt = atoi(env);
    if (nt > 512) nt = 512;
    nt = (nt / 32) * 32;
    if (nt < 32) nt = 32;
    if (nt < TEAM) nt = TEAM;
    while (nt > 64 && (nt / 2) / TEAM >= L) nt /= 2;
    while (nt > 32 && nt > TEAM && amp_smem<LOGM>(n, nt / TEAM) > 227 * 1024) nt /= 2;
    *W_out = nt / TEAM;
    *smem_out = amp_smem<LOGM>(n, *W_out);
    return nt;
}

template <int LOGM, bool PRE>
static int launch_amp(const sb_operator *op, AmpArgs a, int B, int which, const double *in, double *out,
                      cudaStream_t st) {
    size_t smem = 0;
    int W = 0;
    const int nt = pick_threads<LOGM>(op->n, op->L, &smem, &W);
    if (smem > 227 * 1024) return fail(SB_EINVAL, "AMP: n too large for shared memory%s (%ld bytes)", "", (long)smem);
    if (which == 0) {
        SB_CUDA(cudaFuncSetAttribute(amp_kernel<LOGM, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        amp_kernel<LOGM, PRE><<<B, nt, smem, st>>>(a, W);
    } else if (which == 1) {
        SB_CUDA(cudaFuncSetAttribute(Ab_kernel<LOGM, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Ab_kernel<LOGM, PRE><<<B, nt, smem, st>>>(a, W, in, out);
    } else {
        SB_CUDA(cudaFuncSetAttribute(Az_kernel<LOGM, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Az_kernel<LOGM, PRE><<<B, nt, smem, st>>>(a, W, in, out);
    }
    SB_LAUNCHED();
    return SB_OK;
}

static int dispatch(const sb_operator *op, AmpArgs a, int B, int which, const double *in, double *out,
                    cudaStream_t st) {
    switch (op->logM) {
#define SB_CASE(l)                                                                  \
    case l:                                                                         \
        return op->pre ? launch_amp<l, true>(op, a, B, which, in, out, st)          \
                       : launch_amp<l, false>(op, a, B, which, in, out, st);
        SB_CASE(1) SB_CASE(2) SB_CASE(3) SB_CASE(4) SB_CASE(5) SB_CASE(6) SB_CASE(7) SB_CASE(8) SB_CASE(9) SB_CASE(10)
#undef SB_CASE
    }
    return fail(SB_EINVAL, "unsupported section size M = 2^%s%ld", "", op->logM);
}

static int sign_bit(int logM) { return logM + 3 + (logM >= 10 ? 2 : 3); }

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
    const int logH = ilog2(op->H), SB = sign_bit(op->logM);
    op->pre = ((size_t)n * 8 <= 65535) ? 1 : 0;  // inverse-table entries as byte offsets when they fit in 16 bits
    op->G8 = (L + 7) / 8;
    const size_t nf = (size_t)L * n, ni = (size_t)L * M * op->Hp, n8 = (size_t)op->G8 * n * 8;
    uint16_t *hf = (uint16_t *)malloc(nf * 2), *hi = (uint16_t *)malloc(ni * 2), *h8 = (uint16_t *)calloc(n8, 2);
    if (!hf || !hi || !h8) { free(hf); free(hi); free(h8); delete op; return fail(SB_ENOMEM, "sb_operator_create: host alloc%s", ""); }
    const uint16_t empty = (uint16_t)(op->pre ? n * 8 : n);  // the zero word zs[n]
    for (size_t i = 0; i < ni; i++) hi[i] = empty;
    for (int l = 0; l < L; l++)
        for (int k = 0; k < n; k++) {
            const uint32_t r = ordering[(size_t)l * n + k];
            if (r == 0 || r >= (uint32_t)w) { free(hf); free(hi); free(h8); delete op; return fail(SB_EINVAL, "ordering entry out of [1,w)%s", ""); }
            const uint32_t lo = r % M, hiw = r / M;
            uint32_t c = 0;  // visit position = bit reversal of the block index over log2(H) bits
            for (int bbit = 0; bbit < logH; bbit++) c |= ((hiw >> bbit) & 1u) << (logH - 1 - bbit);
            const uint16_t fe = (uint16_t)((lo << 3) | ((__builtin_popcount(hiw) & 1) << SB));
            hf[(size_t)l * n + k] = fe;
            h8[((size_t)(l >> 3) * n + k) * 8 + (l & 7)] = fe;
            hi[((size_t)l * M + lo) * op->Hp + c] = (uint16_t)(op->pre ? k * 8 : k);
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
                            const int *sections, const int *nsec, int B, int T, double *beta, int *iters, int *n_exec,
                            unsigned *flags, double *tau2_trace, void *stream) {
    if (!op || !y || !Pl || !beta || !iters || !n_exec || !flags || B < 0 || T < 0)
        return fail(SB_EINVAL, "sb_amp_batch: null argument%s", "");
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_amp_batch: sections and nsec go together%s", "");
    if (B == 0) return SB_OK;
    AmpArgs a = base_args(op, sections, nsec);
    a.y = y; a.Pl = Pl; a.beta0 = beta0; a.beta = beta; a.tau2_trace = tau2_trace;
    a.iters = iters; a.n_exec = n_exec; a.flags = flags; a.T = T;
    return dispatch(op, a, B, 0, nullptr, nullptr, (cudaStream_t)stream);
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
