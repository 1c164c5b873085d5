// amp_impl.cuh -- SPARC design operator and AMP decoder kernels (sm_100a), templated on the section size.
//
// Replaces ldpc/sparc_ldpc.py:32-147 (sub_fht / block_sub_fht / sparc_transforms[_shorter]) and
// :189-222 (amp) of the reference.  The reference zero-pads every section to w = 2^ceil(log2(n+1))
// and runs a w-point Walsh-Hadamard transform; because only the last M columns of H_w are used,
//     (A beta)[k]  = (1/sqrt n) sum_l sgn(l,k) * FHT_M(beta_l)[lo(l,k)]
//     (A^T z)_l    = (1/sqrt n) FHT_M(fold_l(z)),
// with r = ordering[l][k], lo = r mod M, hi = r div M, sgn = (-1)^popcount(hi) and fold_l the signed
// butterfly tree over the w/M blocks.  In STRICT mode both identities are evaluated in the reference's
// own order of additions (fold tree = order of the large-stride butterflies, M-point stages from stride
// M/2 down to 1, sections accumulated into A beta in ascending order, sparc_ldpc.py:123-126), which makes
// Ab / Az bit-identical to the reference.
//
// Kernel structure: ONE persistent CTA per codeword runs the whole AMP loop (all T iterations) in a
// single launch.  z and the A beta accumulator live in shared memory; beta streams through HBM once in
// and once out per iteration (the section-wise softmax needs only the section itself); the lookup tables
// are shared by the whole batch and stay L2-resident:
//     inv  u16 [L][M][Hp]   for bin lo of section l: the k (as k*4) of every block, in the visit order of the
//                           fold tree; empty blocks point at a zero word
//     fwd  u16 [L][n]       for row k of section l: lo*4 | sgn << SBQ   (section lists)
//     fwd8 u16 [L/8][n][8]  the same entries with 8 sections interleaved: one 16-byte load per (group, k)
// Per group of W sections each team (<= one warp) transforms one section and leaves FHT_M(beta_l) in shared
// memory TWICE, as +F and as -F, so that the signed gather of A beta is one LDS at offset (entry) and one add.
//
// The kernel is bound by the shared-memory data pipe: 2*L*n random accesses per codeword-iteration.  QUANT
// mode therefore keeps 32-bit fixed-point copies of the two randomly gathered vectors: z (27 bits below its
// per-iteration power-of-two ceiling) and F (27 bits below the ceiling of sqrt(n P_l)); fold and gather
// become exact integer adds, everything else (FHT, softmax, z update, tau^2) stays fp64.  The only error is
// the rounding of z and F to 2^-27 of their range (relative effect on beta ~1e-8, tolerance 1e-5).
#pragma once
#include "common.cuh"

namespace sb {

__host__ __device__ __forceinline__ int pad2(int n) { return (n + 1) & ~1; }
__host__ __device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }

template <int LOGM>
struct TeamCfg {
    static constexpr int M = 1 << LOGM;
    static constexpr int TEAM = (M >= 128) ? 32 : (M >= 4 ? M / 4 : 1);  // lanes cooperating on one section
    static constexpr int EPT = M / TEAM;                                  // elements per lane
    static constexpr int SPR = (LOGM >= 10) ? 4 : 8;                      // section slots per F region
    static constexpr int SBQ = LOGM + 2 + ((LOGM >= 10) ? 2 : 3);         // sign bit of a fwd entry (<= 14)
};

// byte offset of section slot `slot` (+F copy) inside the F area; ESH = log2(element bytes) (3: fp64, 2: int32)
template <int LOGM, int ESH>
__device__ __forceinline__ int slot_offset(int slot) {
    using C = TeamCfg<LOGM>;
    return (slot / C::SPR) * (2 << (C::SBQ + ESH - 2)) + (slot % C::SPR) * (C::M << ESH);
}

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

// ---- FAST mode, M = 512: transforms with ONE shuffle stage ----------------------------------------------------
// Two register layouts of the 512 values of a section held by a warp (lane q, register index i):
//   layout A: j = i * 32 + q                              (coalesced beta access; registers = bits 8..5 of j)
//   layout B: j = (q >> 1) * 32 + 2 * i + (q & 1)         (registers = bits 4..1 of j, lane bit 0 = bit 0 of j)
// A transform = 4 register stages in one layout, a transpose through the warp's own (still unused) +-F slot in
// shared memory, 4 register stages in the other layout, and one xor-1 shuffle stage for bit 0 of j: 96
// shared-memory wavefronts instead of the 160 of five shuffle stages, and ~150 fewer instructions.
// Staging position of element (row = j >> 5, col = j & 31), in doubles: row * 32 + (col ^ 2 row), rows 0..7 in
// the +F half of the slot and rows 8..15 in the -F half; the xor keeps both access patterns conflict-free.
__device__ __forceinline__ double *stage_at(int *Fp, int *Fn, int row, int col) {
    double *base = reinterpret_cast<double *>(row < 8 ? Fp : Fn);
    return base + (row & 7) * 32 + (col ^ ((2 * row) & 31));
}

__device__ __forceinline__ void reg_stages16(double (&x)[16]) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if ((i & s) == 0) {
                const double a = x[i], b = x[i + s];
                x[i] = a + b;
                x[i + s] = a - b;
            }
        }
    }
}

__device__ __forceinline__ void shuffle_stage1(double (&x)[16], int q) {
    const int sgn = (q & 1) ? (int)0x80000000 : 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const double p = __shfl_xor_sync(0xffffffffu, x[i], 1);
        const double mine = __hiloint2double(__double2hiint(x[i]) ^ sgn, __double2loint(x[i]));
        x[i] = p + mine;
    }
}

// layout B in, layout A out; stage order 16, 8, 4, 2, 1, 256, 128, 64, 32 (exact for the integer-valued fold)
__device__ __forceinline__ void fht512_B_to_A(double (&x)[16], int q, int *Fp, int *Fn) {
    reg_stages16(x);
    shuffle_stage1(x, q);
    const int a = q >> 1, b = q & 1;
#pragma unroll
    for (int i = 0; i < 16; i++) *stage_at(Fp, Fn, a, 2 * i + b) = x[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = *stage_at(Fp, Fn, i, q);
    __syncwarp();
    reg_stages16(x);
}

// layout A in, layout B out; stage order 256, ..., 2, 1 = the reference's (ldpc/sparc_ldpc.py:19-29)
__device__ __forceinline__ void fht512_A_to_B(double (&x)[16], int q, int *Fp, int *Fn) {
    reg_stages16(x);
#pragma unroll
    for (int i = 0; i < 16; i++) *stage_at(Fp, Fn, i, q) = x[i];
    __syncwarp();
    const int a = q >> 1, b = q & 1;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = *stage_at(Fp, Fn, a, 2 * i + b);
    __syncwarp();
    reg_stages16(x);
    shuffle_stage1(x, q);
}

// ---- FAST mode: the SECOND transform (beta_l -> F_l) in exact 32-bit fixed point ---------------------------------
// beta_l >= 0 sums to sqrt(n P_l) <= cmax, so with beta quantised to 30 bits below the power-of-two ceiling of cmax every
// butterfly value is bounded by sum_j beta_q[j] < 2^31: the integer transform is exact, its transpose moves 4-byte words
// (32 shared-memory wavefronts instead of 64) and its shuffle stage one register per element instead of two.  F is then
// rounded from 30 to the 27 bits the gathers use.  Error of F in units of 2^-27 cmax: 0.8 rms (512 rounded inputs) instead
// of 0.3 (one rounding) -- two orders of magnitude inside FAST's tolerance.  In the alternating-phase kernel of round 1
// this variant lost (the integer pipe was the busier one); with the transform warps on their own it wins.
#ifndef SB_INT_FHT2
#define SB_INT_FHT2 1
#endif
__device__ __forceinline__ int *stage_at_i(int *base, int row, int col) { return base + row * 32 + (col ^ ((2 * row) & 31)); }

__device__ __forceinline__ void reg_stages16_i(int (&x)[16]) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if ((i & s) == 0) {
                const int a = x[i], b = x[i + s];
                x[i] = a + b;
                x[i + s] = a - b;
            }
        }
    }
}

// layout A in, layout B out (as fht512_A_to_B); S: 2 KB of the warp's own shared memory
__device__ __forceinline__ void fht512_A_to_B_i(int (&x)[16], int q, int *S) {
    reg_stages16_i(x);
#pragma unroll
    for (int i = 0; i < 16; i++) *stage_at_i(S, i, q) = x[i];
    __syncwarp();
    const int a = q >> 1, b = q & 1;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = *stage_at_i(S, a, 2 * i + b);
    __syncwarp();
    reg_stages16_i(x);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int p = __shfl_xor_sync(0xffffffffu, x[i], 1);
        x[i] = b ? p - x[i] : p + x[i];
    }
}

// physical word of F_l[lo] inside its +-F copy (FAST mode, M = 512): the transform ends in layout B, and
// p(lo) = ((lo >> 1) & 15) * 32 + (lo >> 5) * 2 + (lo & 1) makes the store of register i a contiguous 128 bytes
__host__ __device__ __forceinline__ uint32_t fq_word(int logM, uint32_t lo) {
    return logM == 9 ? (((lo >> 1) & 15u) * 32u + (lo >> 5) * 2u + (lo & 1u)) : lo;
}

// z value addressed by an inverse-table entry e.  PRE: e = k*4 (byte offset of an int32), else e = k.
template <bool PRE, typename T>
__device__ __forceinline__ T zs_at(const T *zs, uint32_t e) {
    if (PRE) return *reinterpret_cast<const T *>(reinterpret_cast<const char *>(zs) + (sizeof(T) == 8 ? (e << 1) : e));
    return zs[e];
}

// One block of 16 inverse-table entries in visit order; every tree node is (left - right), i.e.
// v <- v[:half] - v[half:] of the reference's large-stride butterflies.  Empty slots point at zs[n] = 0.
template <bool PRE, typename T>
__device__ __forceinline__ T fold16(const uint4 p0, const uint4 p1, const T *zs) {
    const uint32_t wds[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    T v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[2 * i] = zs_at<PRE, T>(zs, wds[i] & 0xFFFFu);
        v[2 * i + 1] = zs_at<PRE, T>(zs, wds[i] >> 16);
    }
#pragma unroll
    for (int s = 1; s < 16; s <<= 1) {
#pragma unroll
        for (int i = 0; i < 16; i += 2 * s) v[i] = v[i] - v[i + s];
    }
    return v[0];
}

// One bin with NB > 1 blocks: binary-counter merge (in fp64) of the block subtrees (left - right at every level).
template <bool PRE, typename T>
__device__ __forceinline__ double fold_bin_multi(const uint16_t *__restrict__ tab, int NB, const T *zs, double unit) {
    double st[8];
    double val = 0.0;
    for (int c = 0; c < NB; c++) {
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(tab + c * 16));
        const uint4 p1 = __ldg(reinterpret_cast<const uint4 *>(tab + c * 16 + 8));
        val = (double)fold16<PRE, T>(p0, p1, zs) * unit;
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

// fold_l(z) for the EPT bins of this lane, in units of `unit` (1.0 for fp64 z, 2^-zshift for fixed-point z).
// NB == 1 (w/M <= 16, the headline shapes): the two 16-byte table loads of bin e+1 are issued before bin e
// is reduced.
template <int LOGM, bool PRE, typename T>
__device__ __forceinline__ void fold_section(double (&x)[TeamCfg<LOGM>::EPT], const uint16_t *__restrict__ tab,
                                             int Hp, int NB, int q, const T *zs, double unit) {
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
            if (sizeof(T) == 8)
                x[e] = (double)fold16<PRE, T>(c0, c1, zs);
            else
                x[e] = (double)fold16<PRE, T>(c0, c1, zs) * unit;
            c0 = n0;
            c1 = n1;
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; e++)
            x[e] = fold_bin_multi<PRE, T>(tab + (size_t)(e * TEAM + q) * Hp, NB, zs, sizeof(T) == 8 ? 1.0 : unit);
    }
}

// FAST mode: the bin's terms are signed fixed-point words (+z or -z copy, or a zero word) in an order chosen at
// table-build time so that the lanes of a warp read distinct banks (sched.h); integer adds are order-free.
template <bool PRE>
__device__ __forceinline__ int fold16q(const uint4 p0, const uint4 p1, const int *zs) {
    const uint32_t wds[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += zs_at<PRE, int>(zs, wds[i] & 0xFFFFu) + zs_at<PRE, int>(zs, wds[i] >> 16);
    return s;
}

// table of one section: [EPT][2*NB][TEAM][8 entries].  NBT = 1 | 2 (w/M <= 16 | 32, the headline and the C5
// shapes): fully unrolled, the two 16-byte loads of the next 16-entry chunk are issued before the current chunk
// is reduced.  NBT = 0: any NB, plain loop.
template <int LOGM, bool PRE, int NBT>
__device__ __forceinline__ void fold_section_q(double (&x)[TeamCfg<LOGM>::EPT], const uint16_t *__restrict__ tab,
                                               int NB, int q, const int *zs, double unit) {
    constexpr int TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(tab) + q;
    if (NBT > 0) {
        constexpr int NCH = EPT * (NBT > 0 ? NBT : 1);  // chunks of 16 entries, bin-major
        uint4 c0 = __ldg(t4), c1 = __ldg(t4 + TEAM);
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            uint4 n0 = c0, n1 = c1;
            if (i + 1 < NCH) {
                n0 = __ldg(t4 + (2 * (i + 1)) * TEAM);
                n1 = __ldg(t4 + (2 * (i + 1) + 1) * TEAM);
            }
            const int s = fold16q<PRE>(c0, c1, zs);
            if (NBT == 1) {
                x[i] = (double)s * unit;
            } else {
                v = (i % NBT == 0) ? (double)s : v + (double)s;
                if (i % NBT == NBT - 1) x[i / NBT] = v * unit;
            }
            c0 = n0;
            c1 = n1;
        }
    } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const uint4 *te = t4 + (size_t)e * 2 * NB * TEAM;
            double v = 0.0;
            for (int c = 0; c < NB; c++)
                v += (double)fold16q<PRE>(__ldg(te + (2 * c) * TEAM), __ldg(te + (2 * c + 1) * TEAM), zs);
            x[e] = v * unit;
        }
    }
}

#ifndef SB_FAST_EXP
#define SB_FAST_EXP 1  // FAST mode: exp(x), x <= 0, with a degree-10 polynomial (rel. error 2e-13) instead of CUDA's exp
#endif
// exp(x) for x <= 0 (the softmax argument minus its section maximum), FAST mode only: n = rint(x log2 e) by the 2^52 + 2^51
// trick, r = x - n ln 2 in two FMAs, e^r by its Taylor polynomial of degree 10 on |r| <= 0.347 (truncation 2.2e-13 relative,
// five orders of magnitude below FAST's quantisation), scaled by adding n to the exponent field.  x < -708 returns 0 (the
// result would be subnormal; CUDA's exp returns the subnormal -- the softmax sum does not see the difference at 1e-308).
// 18 instructions instead of 24 (pair kernel: 2.30 -> 2.28 us per codeword-iteration); both FAST kernels use it, so that
// they stay interchangeable to fp64 summation-order noise; F64 / STRICT keep CUDA's exp.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 2.7557319223985888e-07;             // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);         // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);         // 1/8!
    p = fma(p, r, 1.9841269841269841e-04);         // 1/7!
    p = fma(p, r, 1.3888888888888889e-03);         // 1/6!
    p = fma(p, r, 8.3333333333333332e-03);         // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);         // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);         // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double s = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return (x < -708.0) ? 0.0 : s;
}

struct AmpArgs {
    const uint16_t *fwd, *fwd8, *inv;
    const uint16_t *invq, *fwdq;  // FAST mode: scheduled (bank-conflict-free) orderings of the same maps
    int qneg;                     // word offset of the -z copy in the fixed-point z area
    int PW;                       // sections per scheduled gather chunk (8 | 16)
    const double *y, *Pl, *beta0;
    const int *sections, *nsec;
    double *beta, *tau2_trace, *zscratch;
    int *iters, *n_exec;
    unsigned *flags;
    int L, n, Hp, NB, T;
#ifdef SB_PHASE_CLOCKS
    unsigned long long *dbg;  // experiment builds only: cycles per phase, accumulated by thread 0 of every CTA
#endif
};

// Experiment builds (make EXTRA=-DSB_PHASE_CLOCKS, tools/phase_clocks.py): thread 0 adds the cycles since its
// previous mark to dbg[i].
#ifdef SB_PHASE_CLOCKS
__device__ __forceinline__ unsigned &sb_clk_last() {
    __shared__ unsigned last;
    return last;
}
#define SB_CLK(a, i)                                                             \
    do {                                                                         \
        if (threadIdx.x == 0) {                                                  \
            const unsigned c__ = (unsigned)clock();                              \
            atomicAdd((a).dbg + (i), (unsigned long long)(c__ - sb_clk_last())); \
            sb_clk_last() = c__;                                                 \
        }                                                                        \
    } while (0)
#else
#define SB_CLK(a, i) do { } while (0)
#endif

struct SecCtx {  // per-iteration scalars of the section phase
    double inv_rt_n, tau2, zunit, fscale;
};

// mode 0: AMP iteration (fold -> FHT -> softmax -> store beta -> FHT -> +-F)
// mode 1: operator only (load beta -> FHT -> +-F)   [prologue z = y - A beta0, sb_Ab_batch]
template <int LOGM, bool PRE, bool QUANT, int NBT>
__device__ __forceinline__ void section_phase(int mode, bool first_zero, const AmpArgs &a, const double *bsrc,
                                              double *bdst, int sec, int q, const void *zsv, char *Fbytes, int slot,
                                              const SecCtx &cx, double rt_npl, double &sq, double &gmax, double &lmin) {
    constexpr int M = TeamCfg<LOGM>::M, TEAM = TeamCfg<LOGM>::TEAM, EPT = TeamCfg<LOGM>::EPT;
    constexpr int SBQ = TeamCfg<LOGM>::SBQ;
    constexpr bool TRQ = QUANT && LOGM == 9;  // transposed transforms (one shuffle stage), F stored at fq_word()
    double x[EPT];
    const unsigned tmask = team_mask<TEAM>();
    int *Fqp = reinterpret_cast<int *>(Fbytes + slot_offset<LOGM, 2>(slot));
    int *Fqn = reinterpret_cast<int *>(Fbytes + slot_offset<LOGM, 2>(slot) + (1 << SBQ));
    if (mode == 0) {
        if (QUANT) {  // (for M = 512 the table lists the bins in layout B)
            const uint16_t *tab = a.invq + ((size_t)sec * M) * a.Hp;
            const int *zq = static_cast<const int *>(zsv);
            fold_section_q<LOGM, PRE, NBT>(x, tab, a.NB, q, zq, cx.zunit);
        } else
            fold_section<LOGM, PRE, double>(x, a.inv + ((size_t)sec * M) * a.Hp, a.Hp, a.NB, q,
                                            static_cast<const double *>(zsv), 1.0);
        SB_CLK(a, 1);
        // beta_l streams from HBM: FAST issues its loads before the first transform, which hides their latency
        // (issuing them before the fold as well costs the fold its registers: 3.27 vs 3.09 us measured)
        double bv[TRQ ? EPT : 1];
        if constexpr (TRQ) {
#pragma unroll
            // beta streams through once per iteration: L1::no_allocate loads and streaming stores keep it out of
            // the L1 lines that the table loads in flight need (2.94 -> 2.81 us per codeword-iteration)
            // (the table loads themselves want the default policy: L1::no_allocate / evict_first on them cost 4 %)
            for (int e = 0; e < EPT; e++) {
                bv[e] = 0.0;
                if (!first_zero) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(bv[e]) : "l"(bsrc + e * TEAM + q));
            }
            fht512_B_to_A(x, q, Fqp, Fqn);
        } else {
            fht_team<LOGM>(x, q, tmask);
        }
        SB_CLK(a, 2);
        const double c2 = rt_npl / cx.tau2;
        double m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            double b;
            if constexpr (TRQ) b = bv[e];
            else b = first_zero ? 0.0 : __ldcs(bsrc + e * TEAM + q);  // streamed once per iteration
            const double s = b + x[e] * cx.inv_rt_n;  // s = beta + A^T z          (sparc_ldpc.py:213)
            x[e] = s * c2;                            // u = s sqrt(n P_l)/tau^2    (:215)
            m = fmax(m, x[e]);
        }
#pragma unroll
        for (int d = TEAM / 2; d >= 1; d >>= 1) m = fmax(m, __shfl_xor_sync(tmask, m, d));
        gmax = fmax(gmax, m);
        lmin = fmin(lmin, m);
        double sum = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            // section max instead of the reference's global max (:216): same softmax
            x[e] = (QUANT && SB_FAST_EXP) ? exp_nonpos(x[e] - m) : exp(x[e] - m);
            sum += x[e];
        }
#pragma unroll
        for (int d = TEAM / 2; d >= 1; d >>= 1) sum += __shfl_xor_sync(tmask, sum, d);
        const double sc = rt_npl / sum;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            x[e] *= sc;  // beta = sqrt(n P_l) softmax(u)    (:218-219)
            sq += x[e] * x[e];
            __stcs(bdst + e * TEAM + q, x[e]);
        }
        SB_CLK(a, 3);
    } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            x[e] = bsrc[e * TEAM + q];
            if (bdst != nullptr) bdst[e * TEAM + q] = x[e];
        }
    }
    if constexpr (TRQ) {
#if SB_INT_FHT2
        int xi[EPT];
        const double fs30 = cx.fscale * 8.0;
#pragma unroll
        for (int e = 0; e < EPT; e++) xi[e] = __double2int_rn(x[e] * fs30);
        fht512_A_to_B_i(xi, q, Fqp);
#pragma unroll
        for (int e = 0; e < EPT; e++) {  // register e of lane q holds lo = (q>>1)*32 + 2e + (q&1): word e*32 + q
            const int f = (xi[e] + 4) >> 3;
            Fqp[e * 32 + q] = f;
            Fqn[e * 32 + q] = -f;
        }
#else
        fht512_A_to_B(x, q, Fqp, Fqn);
#pragma unroll
        for (int e = 0; e < EPT; e++) {  // register e of lane q holds lo = (q>>1)*32 + 2e + (q&1): word e*32 + q
            const int f = __double2int_rn(x[e] * cx.fscale);
            Fqp[e * 32 + q] = f;
            Fqn[e * 32 + q] = -f;
        }
#endif
        SB_CLK(a, 4);
        return;
    }
    fht_team<LOGM>(x, q, tmask);
    if (QUANT) {
        int *Fp = Fqp, *Fn = Fqn;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const int f = __double2int_rn(x[e] * cx.fscale);
            Fp[e * TEAM + q] = f;
            Fn[e * TEAM + q] = -f;
        }
    } else {
        double *Fp = reinterpret_cast<double *>(Fbytes + slot_offset<LOGM, 3>(slot));
        double *Fn = reinterpret_cast<double *>(Fbytes + slot_offset<LOGM, 3>(slot) + (2 << SBQ));
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            Fp[e * TEAM + q] = x[e];
            Fn[e * TEAM + q] = -x[e];
        }
    }
}

// signed F value addressed by a fwd entry (lo*4 | sgn << SBQ) relative to the slot's +F copy
template <typename T>
__device__ __forceinline__ T F_at(const char *slot_base, uint32_t e) {
    return *reinterpret_cast<const T *>(slot_base + (sizeof(T) == 8 ? (e << 1) : e));
}

// acc[k] += sum over the group's sections (ascending) of sgn * F[lo]     (sparc_ldpc.py:123-126, :70)
// generic section lists: one u16 row per section.  T = double (strict) or int (fixed point, flushed to fp64
// every 16 sections so that the int32 partial sums cannot overflow).
// fwd entry (lo*4 | sgn << SBQ) -> offset of the signed F word in the layout of the FAST M = 512 slot
template <int LOGM>
__device__ __forceinline__ uint32_t fq_entry(uint32_t e) {
    constexpr int SBQ = TeamCfg<LOGM>::SBQ;
    if (LOGM != 9) return e;
    return (fq_word(9, (e & ((1u << SBQ) - 1u)) >> 2) << 2) | (e & (1u << SBQ));
}

template <int LOGM, typename T>
__device__ __forceinline__ void gather_phase(const uint16_t *__restrict__ fwd, int n, int nvalid, const int *sec_s,
                                             const char *Fbytes, double *acc_s, double funit, int first = 0) {
    constexpr int ESH = sizeof(T) == 8 ? 3 : 2;
    constexpr bool REMAP = sizeof(T) == 4 && LOGM == 9;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double acc = acc_s[k];
        T part = 0;
        int tm = first;
        for (; tm + 4 <= nvalid; tm += 4) {
            uint32_t e[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                e[i] = __ldg(fwd + (size_t)sec_s[tm + i] * n + k);
                if (REMAP) e[i] = fq_entry<LOGM>(e[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const T v = F_at<T>(Fbytes + slot_offset<LOGM, ESH>(tm + i), e[i]);
                if (sizeof(T) == 8) acc += v; else part += v;
            }
            if (sizeof(T) != 8 && ((tm + 4 - first) & 15) == 0) {
                acc += (double)part * funit;
                part = 0;
            }
        }
        for (; tm < nvalid; tm++) {
            uint32_t e1 = __ldg(fwd + (size_t)sec_s[tm] * n + k);
            if (REMAP) e1 = fq_entry<LOGM>(e1);
            const T v = F_at<T>(Fbytes + slot_offset<LOGM, ESH>(tm), e1);
            if (sizeof(T) == 8) acc += v; else part += v;
        }
        if (sizeof(T) != 8) acc += (double)part * funit;
        acc_s[k] = acc;
    }
}

// all sections in order and groups aligned to 8: the 8 entries of (group, k) are one 16-byte load from the
// interleaved table.  KB rows per thread are in flight together.  Full pairs of chunks (16 sections, the
// headline group size) take a predicate-free path; a trailing chunk takes the guarded one.
template <int LOGM, typename T, int KB>
__device__ __forceinline__ void gather_chunk(const uint4 (&w)[KB], const char *F, int nv, double (&acc)[KB],
                                             T (&part)[KB]) {
    constexpr int ESH = sizeof(T) == 8 ? 3 : 2;
#pragma unroll
    for (int j = 0; j < KB; j++) {
        const uint32_t wds[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (nv >= 8 || i < nv) {
                const uint32_t e = (i & 1) ? (wds[i >> 1] >> 16) : (wds[i >> 1] & 0xFFFFu);
                const T v = F_at<T>(F + slot_offset<LOGM, ESH>(i), e);
                if (sizeof(T) == 8) acc[j] += v; else part[j] += v;
            }
        }
    }
}

template <int LOGM, typename T>
__device__ __forceinline__ void gather_phase8(const uint16_t *__restrict__ fwd8, int n, int g0, int nvalid,
                                              const char *Fbytes, double *acc_s, double funit) {
    constexpr int KB = 3;
    constexpr int ESH = sizeof(T) == 8 ? 3 : 2;
    const uint4 *tab = reinterpret_cast<const uint4 *>(fwd8) + (size_t)(g0 >> 3) * n;
    const int nfull = nvalid >> 3, NT = blockDim.x;  // complete chunks of 8 sections
    for (int k0 = threadIdx.x; k0 < n; k0 += KB * NT) {
        double acc[KB];
        T part[KB];
        int kk[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) {
            const int k = k0 + j * NT;
            kk[j] = (k < n) ? k : k0;  // clamp: a duplicate row is gathered and then discarded
            acc[j] = acc_s[kk[j]];
            part[j] = 0;
        }
        int c = 0;
        for (; c + 2 <= nfull; c += 2) {  // 16 sections: two 16-byte loads per row, no predicates
            uint4 w0[KB], w1[KB];
#pragma unroll
            for (int j = 0; j < KB; j++) {
                w0[j] = __ldg(tab + (size_t)c * n + kk[j]);
                w1[j] = __ldg(tab + (size_t)(c + 1) * n + kk[j]);
            }
            const char *F = Fbytes + slot_offset<LOGM, ESH>(c * 8);
            gather_chunk<LOGM, T, KB>(w0, F, 8, acc, part);
            gather_chunk<LOGM, T, KB>(w1, F + slot_offset<LOGM, ESH>(8), 8, acc, part);
            if (sizeof(T) != 8) {  // <= 16 int32 terms of < 2^27 each per flush
#pragma unroll
                for (int j = 0; j < KB; j++) {
                    acc[j] += (double)part[j] * funit;
                    part[j] = 0;
                }
            }
        }
        for (; c * 8 < nvalid; c++) {  // trailing chunk(s), possibly partial
            uint4 w0[KB];
#pragma unroll
            for (int j = 0; j < KB; j++) w0[j] = __ldg(tab + (size_t)c * n + kk[j]);
            gather_chunk<LOGM, T, KB>(w0, Fbytes + slot_offset<LOGM, ESH>(c * 8), nvalid - c * 8, acc, part);
            if (sizeof(T) != 8) {
#pragma unroll
                for (int j = 0; j < KB; j++) {
                    acc[j] += (double)part[j] * funit;
                    part[j] = 0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KB; j++) {
            const int k = k0 + j * NT;
            if (k < n) acc_s[k] = acc[j];
        }
    }
}

// FAST mode, all sections in order: the PW entries of (PW-section chunk, row k) are PW/8 16-byte loads from the
// scheduled table [chunk][PW/8][n][8]; every entry is the byte offset of its signed F word inside the chunk's
// +-F area, and the order of the entries differs from row to row such that the 32 rows of a warp read distinct
// banks (sched.h).
template <int LOGM, int PW>
__device__ __forceinline__ void gather_phaseq(const uint16_t *__restrict__ fwdq, int n, int g0, int nchunks,
                                              const char *Fbytes, double *acc_s, double funit) {
    constexpr int KB = 3, NH = PW / 8;
    const uint4 *tab = reinterpret_cast<const uint4 *>(fwdq) + (size_t)(g0 / PW) * NH * n;
    const int NT = blockDim.x;
    for (int k0 = threadIdx.x; k0 < n; k0 += KB * NT) {
        double acc[KB];
        int kk[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) {
            const int k = k0 + j * NT;
            kk[j] = (k < n) ? k : k0;  // clamp: a duplicate row is gathered and then discarded
            acc[j] = acc_s[kk[j]];
        }
        for (int c = 0; c < nchunks; c += 16 / PW) {  // <= 16 terms of < 2^27 each per int32 partial sum
            uint4 w[16 / PW][NH][KB];
#pragma unroll
            for (int cc = 0; cc < 16 / PW; cc++)
#pragma unroll
                for (int h = 0; h < NH; h++)
#pragma unroll
                    for (int j = 0; j < KB; j++)
                        w[cc][h][j] = (c + cc < nchunks) ? __ldg(tab + ((size_t)(c + cc) * NH + h) * n + kk[j])
                                                         : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < KB; j++) {
                int part = 0;
#pragma unroll
                for (int cc = 0; cc < 16 / PW; cc++) {
                    if (c + cc < nchunks) {
                        const char *F = Fbytes + slot_offset<LOGM, 2>((c + cc) * PW);
#pragma unroll
                        for (int h = 0; h < NH; h++) {
                            const uint32_t wds[4] = {w[cc][h][j].x, w[cc][h][j].y, w[cc][h][j].z, w[cc][h][j].w};
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                part += *reinterpret_cast<const int *>(F + (wds[i] & 0xFFFFu)) +
                                        *reinterpret_cast<const int *>(F + (wds[i] >> 16));
                        }
                    }
                }
                acc[j] += (double)part * funit;
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
// Measured and rejected on this structure (tools/ab_build.sh, tools/profile_amp.py; 2.94 us per codeword-iteration
// as is): the first gather-table loads of a group issued before the barrier that publishes F (3.26 us), the row
// batches of the gather software-pipelined through two register buffers (spills, 4.7 us), a second +-F area with
// one barrier per group instead of two (3.31 us), the beta loads moved behind the first transform's transpose (3.09
// us), the second transform in 32-bit fixed point (3.20 us), a deeper table prefetch in the fold (neutral).  The
// kernel sits at the 128-register limit; each of these adds live state that ptxas pays for elsewhere.
// Also rejected: beta through per-warp TMA staging rows (cp.async.bulk + mbarrier; correct, 32 registers freed, HBM
// latency off the warp) -- 3.47 us, because the extra 64 KB of shared memory pushes the carve-out from 164 to 228 KB
// and the kernel needs its L1: the table loads in flight (16 x 512 bytes per warp) live in L1 line buffers.  The
// same kernel at carve-out 196 / 228 KB (32 KB / no L1) runs at 3.33 us instead of 2.94; 96 KB of L1 instead of 64
// brings < 1 % (SB_AMP_CARVEOUT experiments).  This is also why two CTAs per SM (176 KB) lost.
// Gather with the two table halves of a row batch refilled for the next batch as soon as each is consumed (same 24
// registers, order forced through a run-time-zero data dependency because ptxas sinks the refills): 2.92 us --
// the waits it removes were not on the critical path, the phase is bound by the L1 data pipe.
template <int LOGM, bool PRE, bool QUANT, int NBT>
__device__ __forceinline__ void operator_pass(int mode, bool first_zero, const AmpArgs &a, const double *bsrc,
                                              double *bdst, const int *act, int La, const void *zsv, double *acc_s,
                                              char *Fbytes, int *sec_s, int W, const SecCtx &cx, const double *rtp, double &sq,
                                              double &gmax, double &lmin) {
    constexpr int M = TeamCfg<LOGM>::M, TEAM = TeamCfg<LOGM>::TEAM;
    const int tm = threadIdx.x / TEAM, q = threadIdx.x % TEAM;
    const bool qpass = QUANT;
    for (int k = threadIdx.x; k < a.n; k += blockDim.x) acc_s[k] = 0.0;
    for (int g0 = 0; g0 < La; g0 += W) {
        SB_CLK(a, 0);
        const int sidx = g0 + tm;
        const bool valid = (tm < W) && (sidx < La);
        const int sec = valid ? (act ? act[sidx] : sidx) : 0;
        if (q == 0 && tm < W) sec_s[tm] = sec;
        if (valid) {
            const double rt_npl = (mode == 0) ? rtp[sidx] : 0.0;
            section_phase<LOGM, PRE, QUANT, NBT>(mode, first_zero, a, bsrc ? bsrc + (size_t)sidx * M : nullptr,
                                            bdst ? bdst + (size_t)sidx * M : nullptr, sec, q, zsv, Fbytes, tm, cx,
                                            rt_npl, sq, gmax, lmin);
        }
        __syncthreads();
        SB_CLK(a, 5);
        const int nvalid = min(W, La - g0);
        const bool fast = (act == nullptr && (W & 7) == 0);
        const double funit = 1.0 / cx.fscale;
        if (qpass) {
            if (LOGM <= 9 && act == nullptr && a.fwdq != nullptr && nvalid >= a.PW && W % a.PW == 0) {
                const int nch = nvalid / a.PW;  // complete chunks: scheduled table
                if (a.PW == 8) gather_phaseq<LOGM, 8>(a.fwdq, a.n, g0, nch, Fbytes, acc_s, funit);
                else gather_phaseq<LOGM, 16>(a.fwdq, a.n, g0, nch, Fbytes, acc_s, funit);
                if (nch * a.PW < nvalid)
                    gather_phase<LOGM, int>(a.fwd, a.n, nvalid, sec_s, Fbytes, acc_s, funit, nch * a.PW);
            } else if (fast && LOGM != 9) gather_phase8<LOGM, int>(a.fwd8, a.n, g0, nvalid, Fbytes, acc_s, funit);
            else gather_phase<LOGM, int>(a.fwd, a.n, nvalid, sec_s, Fbytes, acc_s, funit);
        } else {
            if (fast) gather_phase8<LOGM, double>(a.fwd8, a.n, g0, nvalid, Fbytes, acc_s, 1.0);
            else gather_phase<LOGM, double>(a.fwd, a.n, nvalid, sec_s, Fbytes, acc_s, 1.0);
        }
        SB_CLK(a, 6);
        __syncthreads();
        SB_CLK(a, 7);
    }
}

// shared-memory carve-up.  strict: z fp64 [n+1 (zero word)] | acc | F (fp64 +-) | red | sec
//                          quant : zq int32 [+z (n) | 32 zero words | -z (n)] | acc | F (int32 +-) | red | sec
//                                  (fp64 z lives in a per-codeword global scratch row: it is only touched
//                                  element-wise)
template <int LOGM, bool QUANT>
struct Smem {
    double *zf, *acc, *red;
    int *zq;
    char *F;
    int *sec;
    __host__ __device__ static size_t f_bytes(int W) {
        const int regions = (W + TeamCfg<LOGM>::SPR - 1) / TeamCfg<LOGM>::SPR;
        return (size_t)regions * ((QUANT ? 2 : 4) << TeamCfg<LOGM>::SBQ);
    }
    // accg: the A beta accumulator lives in the codeword's global scratch row instead (shapes with w/M = 32, i.e.
    // n > 8191: with it in shared memory the CTA needs 203 KB, the carve-out becomes 228 KB and no L1 is left for the
    // table loads in flight -- see the carve-out experiments in DESIGN.md section 6)
    __host__ __device__ static size_t bytes(int n, int W, int accg = 0) {
        size_t b = QUANT ? sizeof(int) * (size_t)pad4(2 * n + 32) : sizeof(double) * (size_t)pad2(n + 1);
        if (!accg) b += sizeof(double) * (size_t)pad2(n);  // acc
        b += f_bytes(W);
        b += sizeof(double) * 40 + sizeof(int) * (size_t)(W + 2);
        return b;
    }
    __device__ Smem(unsigned char *raw, int n, int W, int accg = 0) {
        if (QUANT) {
            zq = reinterpret_cast<int *>(raw);
            zf = nullptr;
            acc = reinterpret_cast<double *>(zq + pad4(2 * n + 32));
        } else {
            zq = nullptr;
            zf = reinterpret_cast<double *>(raw);
            acc = zf + pad2(n + 1);
        }
        F = reinterpret_cast<char *>(acc + (accg ? 0 : pad2(n)));
        red = reinterpret_cast<double *>(F + f_bytes(W));
        sec = reinterpret_cast<int *>(red + 40);
    }
};

// power-of-two ceiling exponent: smallest e with |v| < 2^e  (v > 0 finite)
__device__ __forceinline__ int ceil_exp(double v) { return (v > 0.0) ? ilogb(v) + 1 : 0; }

// NBT: 16-entry chunks per bin of the FAST fold table known at compile time (1 | 2), 0 = read a.NB (and STRICT)
template <int LOGM, bool PRE, bool QUANT, int NBT>
__global__ void __launch_bounds__(512, 1) amp_kernel(AmpArgs a, int W) {
    constexpr int M = TeamCfg<LOGM>::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, b = blockIdx.x;
    constexpr int ACCG = (QUANT && NBT == 2) ? 1 : 0;
    Smem<LOGM, QUANT> sm(smem_raw, n, W, ACCG);
    double *zf = QUANT ? a.zscratch + (size_t)b * n : sm.zf;  // every thread only touches its own k = tid + j*NT
    // (likewise the accumulator: thread tid owns rows tid + j*NT in the gather, the reset and the z update)
    double *acc_s = ACCG ? a.zscratch + ((size_t)gridDim.x + b) * n : sm.acc, *red = sm.red;
    // sqrt(n P_l) of the i-th active section, staged once behind the other arrays (a per-section global load and
    // sqrt at the top of every section stalled the warp: 2.4 % of the stall samples)
    double *rtp = reinterpret_cast<double *>(smem_raw + ((Smem<LOGM, QUANT>::bytes(n, W, ACCG) + 7) & ~(size_t)7));
    if (QUANT) {
        if (threadIdx.x < 32) sm.zq[n + threadIdx.x] = 0;  // one zero word per bank
    } else if (threadIdx.x == 0) {
        zf[n] = 0.0;
    }
    const void *zsv = QUANT ? static_cast<const void *>(sm.zq) : static_cast<const void *>(zf);
#ifdef SB_PHASE_CLOCKS
    if (threadIdx.x == 0) sb_clk_last() = (unsigned)clock();  // (the first mark of a CTA would otherwise add garbage)
#endif

    const int La = a.nsec ? a.nsec[b] : a.L;
    const int *act = a.sections ? a.sections + (size_t)b * a.L : nullptr;
    const double *y = a.y + (size_t)b * n;
    double *beta = a.beta + (size_t)b * a.L * M;
    const double nd = (double)n;
    const double rt_n = sqrt(nd);
    double sq = 0.0, gmax = -INFINITY, lmin = INFINITY;
    SecCtx cx;
    cx.inv_rt_n = 1.0 / rt_n;
    cx.tau2 = 1.0;
    cx.zunit = 1.0;
    cx.fscale = 1.0;

    if (La <= 0) {  // the reference never calls amp() on an empty section set (sparc_ldpc.py:1015)
        if (threadIdx.x == 0) {
            a.iters[b] = 0;
            a.n_exec[b] = 0;
            a.flags[b] = 0;
        }
        return;
    }

    // P = sum of the active sections' power (np.sum(Pl), sparc_ldpc.py:190); cmax bounds |FHT_M(beta_l)|
    double pl = 0.0, plmax = 0.0;
    for (int i = threadIdx.x; i < La; i += blockDim.x) {
        const double p = a.Pl[act ? act[i] : i];
        pl += p;
        plmax = fmax(plmax, p);
        rtp[i] = sqrt(nd * p);
    }
    const double P = block_sum(pl, red);
    const double cmax = sqrt(nd * block_max(plmax, red));
    // |F| <= sqrt(n P_l) <= cmax; the 1e-6 margin keeps |F_q| strictly below 2^27 so 16-term int32 sums cannot overflow
    const double fscale_q = scalbn(1.0, 27 - ceil_exp(cmax * (1.0 + 1e-6)));

    if (a.beta0 != nullptr) {  // z = y - A beta0   (sparc_ldpc.py:197-198)
        const double *b0 = a.beta0 + (size_t)b * a.L * M;
        if (QUANT) {  // |FHT_M(beta0_l)| <= sum_j |beta0_l[j]|: one streaming pass gives the fixed-point scale
            constexpr int TEAM = TeamCfg<LOGM>::TEAM;
            const unsigned tmask = team_mask<TEAM>();
            const int tm = threadIdx.x / TEAM, q = threadIdx.x % TEAM;
            double bound = 0.0;
            for (int sidx = tm; sidx < La; sidx += blockDim.x / TEAM) {
                double s1 = 0.0;
                for (int j = q; j < M; j += TEAM) s1 += fabs(b0[(size_t)sidx * M + j]);
#pragma unroll
                for (int d = TEAM / 2; d >= 1; d >>= 1) s1 += __shfl_xor_sync(tmask, s1, d);
                bound = fmax(bound, s1);
            }
            cx.fscale = scalbn(1.0, 27 - ceil_exp(block_max(bound, red) * (1.0 + 1e-6)));
        }
        operator_pass<LOGM, PRE, QUANT, NBT>(1, false, a, b0, beta, act, La, zsv, acc_s, sm.F, sm.sec, W, cx, rtp, sq, gmax,
                                        lmin);
        for (int k = threadIdx.x; k < n; k += blockDim.x) zf[k] = y[k] - acc_s[k] / rt_n;
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) zf[k] = y[k];
    }
    __syncthreads();

    bool first_zero = (a.beta0 == nullptr);
    double last_tau = 0.0;
    unsigned flags = 0;
    int t = 0, executed = 0;
    for (t = 0; t < a.T; t++) {
        double part = 0.0, zmax = 0.0;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            part += zf[k] * zf[k];
            zmax = fmax(zmax, fabs(zf[k]));
        }
        const double tau = sqrt(block_sum(part, red) / nd);  // (:203)
        // exact-equality stop (:204).  With fixed-point gathers tau jitters at the quantisation floor instead
        // of reaching an exact fp64 fixed point, so FAST mode stops once tau moves by less than 2^-27 relative:
        // beta is then within ~1e-8 of the fixed point the reference iterates on to.
        if (tau == last_tau || (QUANT && fabs(tau - last_tau) <= tau * 7.450580596923828e-09)) {
            flags |= SB_AMP_STOPPED;
            break;
        }
        last_tau = tau;
        cx.tau2 = tau * tau;
        if (a.tau2_trace != nullptr && threadIdx.x == 0) a.tau2_trace[(size_t)b * a.T + t] = cx.tau2;
        if (QUANT) {  // fixed-point copy of z: 27 bits below the power-of-two ceiling of max |z|
            const int ez = ceil_exp(block_max(zmax, red) * (1.0 + 1e-6));
            const double zscale = scalbn(1.0, 27 - ez);
            cx.zunit = scalbn(1.0, ez - 27);
            cx.fscale = fscale_q;
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                const int v = __double2int_rn(zf[k] * zscale);
                sm.zq[k] = v;
                sm.zq[a.qneg + k] = -v;
            }
            __syncthreads();
        }
        sq = 0.0;
        gmax = -INFINITY;
        lmin = INFINITY;
        SB_CLK(a, 8);
        operator_pass<LOGM, PRE, QUANT, NBT>(0, first_zero, a, beta, beta, act, La, zsv, acc_s, sm.F, sm.sec, W, cx, rtp, sq,
                                        gmax, lmin);
        first_zero = false;
        const double sumsq = block_sum(sq, red);
        const double gm = block_max(gmax, red);
        const double lm = -block_max(-lmin, red);
        // the reference subtracts the GLOBAL maximum (:216): a section whose maximum lies more than -log(DBL_MIN) =
        // 708.4 below it is summed in subnormals (relative error up to ~1e-5 observed) and, beyond 745.1, is 0/0
        if (gm - lm > 708.39) flags |= SB_AMP_REF_NAN;
        const double ons = P - sumsq / nd;  // (:220)
        for (int k = threadIdx.x; k < n; k += blockDim.x) zf[k] = (y[k] - acc_s[k] / rt_n) + (zf[k] / cx.tau2) * ons;
        __syncthreads();
        SB_CLK(a, 9);
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

// A_S beta for a batch (sparc_ldpc.py:143-144): out = acc / sqrt(n)   (always strict fp64)
template <int LOGM, bool PRE>
__global__ void __launch_bounds__(512, 1) Ab_kernel(AmpArgs a, int W, const double *beta_in, double *out) {
    constexpr int M = TeamCfg<LOGM>::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, b = blockIdx.x;
    Smem<LOGM, false> sm(smem_raw, n, W);
    const int La = a.nsec ? a.nsec[b] : a.L;
    const int *act = a.sections ? a.sections + (size_t)b * a.L : nullptr;
    double sq = 0, gmax = 0, lmin = 0;
    const double nd = (double)n, rt_n = sqrt(nd);
    SecCtx cx;
    cx.inv_rt_n = 1.0 / rt_n; cx.tau2 = 1.0; cx.zunit = 1.0; cx.fscale = 1.0;
    operator_pass<LOGM, PRE, false, 0>(1, false, a, beta_in + (size_t)b * a.L * M, nullptr, act, La, sm.zf, sm.acc, sm.F,
                                    sm.sec, W, cx, nullptr, sq, gmax, lmin);
    for (int k = threadIdx.x; k < n; k += blockDim.x) out[(size_t)b * n + k] = sm.acc[k] / rt_n;
}

// A_S^T z for a batch (sparc_ldpc.py:145-146): one team per section   (always strict fp64)
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
        fold_section<LOGM, PRE, double>(x, a.inv + ((size_t)sec * M) * a.Hp, a.Hp, a.NB, q, zs, 1.0);
        fht_team<LOGM>(x, q, team_mask<TEAM>());
#pragma unroll
        for (int e = 0; e < EPT; e++) out[(size_t)b * a.L * M + (size_t)sidx * M + e * TEAM + q] = x[e] / rt_n;
    }
}

// threads per CTA.  Default: 512 (one CTA per SM); with fixed-point gathers the state is small enough for TWO
// 256-thread CTAs per SM, whose gather (shared-memory pipe) and transform (fp64 / issue) phases then overlap.
// Never more teams than sections, never more shared memory than 227 KB.
template <int LOGM, bool QUANT>
static int pick_threads(int n, int L, int pw, size_t *smem_out, int *W_out, int accg = 0) {
    constexpr int TEAM = TeamCfg<LOGM>::TEAM;
    int nt = 512;
    // FAST: two 256-thread CTAs per SM when the operator's gather table was built for 8-section chunks (or, for
    // M = 1024, which has no scheduled gather table, whenever two CTAs fit)
    if (QUANT && TEAM == 32 && (pw == 8 || (LOGM > 9 && 2 * (Smem<LOGM, QUANT>::bytes(n, 256 / TEAM) + 1024) <= 227 * 1024)))
        nt = 256;
    // FAST, sections smaller than a warp (M < 128): two 256-thread CTAs per SM, so that one codeword's transform phase
    // overlaps another's gather (measured with the knob below, batch 9472: C4 L=256 M=32 +4 %, C1 L=128 M=4 +10 %)
    if (QUANT && TEAM < 32) nt = 256;
    const char *env = knob("SB_AMP_THREADS");
    if (env) nt = atoi(env);
    if (nt > 512) nt = 512;
    nt = (nt / 32) * 32;
    if (nt < 32) nt = 32;
    if (nt < TEAM) nt = TEAM;
    while (nt > 64 && (nt / 2) / TEAM >= L) nt /= 2;
    while (nt > 32 && nt > TEAM && Smem<LOGM, QUANT>::bytes(n, nt / TEAM, accg) + 8 * (size_t)L + 8 > 227 * 1024) nt /= 2;
    *W_out = nt / TEAM;
    *smem_out = ((Smem<LOGM, QUANT>::bytes(n, *W_out, accg) + 7) & ~(size_t)7) + sizeof(double) * (size_t)L;  // + rtp[L]
    return nt;
}

// which: 0 = AMP strict, 1 = A beta, 2 = A^T z, 3 = AMP with fixed-point gathers
template <int LOGM>
int launch_amp(const sb_operator *op, AmpArgs a, int B, int which, const double *in, double *out, cudaStream_t st) {
    size_t smem = 0;
    int W = 0;
    const bool quant = (which == 3) && op->qok;
    const int nbt0 = (op->NB == 1) ? 1 : ((op->NB == 2 && LOGM == 9) ? 2 : 0);
    const int accg = (quant && nbt0 == 2 && !op->qpre) ? 1 : 0;  // = ACCG of the kernel dispatched below
    const int nt = quant ? pick_threads<LOGM, true>(op->n, op->L, op->PW, &smem, &W, accg)
                         : pick_threads<LOGM, false>(op->n, op->L, 0, &smem, &W);
    if (smem > 227 * 1024) return fail(SB_EINVAL, "AMP: n too large for shared memory%s (%ld bytes)", "", (long)smem);
#define SB_LAUNCH(KERNEL, ...)                                                                           \
    do {                                                                                                 \
        SB_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        KERNEL<<<B, nt, smem, st>>>(__VA_ARGS__);                                                        \
    } while (0)
    if (which == 0 || which == 3) {
        const int nbt = (op->NB == 1) ? 1 : ((op->NB == 2 && LOGM == 9) ? 2 : 0);
        if (quant && op->qpre && nbt == 1) SB_LAUNCH((amp_kernel<LOGM, true, true, 1>), a, W);
        else if (quant && op->qpre) SB_LAUNCH((amp_kernel<LOGM, true, true, 0>), a, W);
        else if (quant && nbt == 1) SB_LAUNCH((amp_kernel<LOGM, false, true, 1>), a, W);
        else if (quant && nbt == 2) SB_LAUNCH((amp_kernel<LOGM, false, true, (LOGM == 9 ? 2 : 0)>), a, W);
        else if (quant) SB_LAUNCH((amp_kernel<LOGM, false, true, 0>), a, W);
        else if (op->pre) SB_LAUNCH((amp_kernel<LOGM, true, false, 0>), a, W);
        else SB_LAUNCH((amp_kernel<LOGM, false, false, 0>), a, W);
    } else if (which == 1) {
        if (op->pre) SB_LAUNCH((Ab_kernel<LOGM, true>), a, W, in, out);
        else SB_LAUNCH((Ab_kernel<LOGM, false>), a, W, in, out);
    } else {
        if (op->pre) SB_LAUNCH((Az_kernel<LOGM, true>), a, W, in, out);
        else SB_LAUNCH((Az_kernel<LOGM, false>), a, W, in, out);
    }
#undef SB_LAUNCH
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
