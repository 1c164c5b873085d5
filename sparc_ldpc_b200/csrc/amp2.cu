// amp2.cu -- AMP decoder for M = 512, warp-specialised (sm_100a): FAST mode with two codewords per CTA
// (amp2_kernel<false>) and F64 mode with one fp64 codeword per CTA (amp2_kernel<true>, SB_AMP_F64).
//
// Same arithmetic as amp_kernel<9, *, QUANT> of amp_impl.cuh (ldpc/sparc_ldpc.py:189-222 over the operator of
// :32-147; fp64 transforms / softmax / z update / tau^2, 27-bit fixed-point gathers), re-organised around what
// bounds that kernel: the L1 / shared-memory data pipe (random fold and gather reads + the L2-resident table
// words) and, in separate phases, the fp64 pipe.  Three changes:
//
//  1. A CTA decodes TWO codewords ("slots").  The z copies and the transformed sections of both codewords sit at
//     the same offset of two planes, so every 16-bit table word is loaded and decoded ONCE for two gathers: the
//     table stream through L1 (25 KB per section, twice the cost of the shared-memory reads it steers because a
//     global load occupies the data pipe two cycles per 128 bytes) is halved per codeword.
//  2. The 16 warps are split by ROLE.  Warps 0-7 ("operator warps") do nothing but fold z into the bins of the
//     next 8 sections and gather the transformed previous 8 sections into A beta -- the data-pipe-bound work;
//     warps 8-15 ("transform warps") run FHT -> softmax -> FHT of the current 8 sections of both codewords -- the
//     fp64-bound work.  Group g+1 is folded and group g-1 gathered WHILE group g is transformed (two 32 KB group
//     buffers, one CTA barrier per group), so the two pipes work concurrently instead of in alternating phases.
//  3. The A beta accumulator lives in the operator warps' REGISTERS (18 rows x 2 codewords of int64 fixed point per
//     thread; the sums of 512 27-bit terms are exact), not in shared memory: no accumulator traffic on the data
//     pipe and 72 KB less shared memory, which is what makes room for two codewords.
//
// No -z / -F copies: the sign of a fold term is the half of the bin's 16 steps it is listed in, the sign of a
// gather term is bit 15 of its table entry (one PRMT + one XOR per codeword on the integer pipes).
//
// Slots are refilled from a per-launch work counter as soon as a codeword stops, so a codeword that runs all T
// iterations never idles its partner.  Integer gathers are order-free, hence A^T z and A beta are bit-identical to
// amp_kernel's FAST path; tau^2 and |beta|^2 are summed over a different thread partition (last-bit differences).
// Both FAST kernels run the second transform (beta_l -> F_l) in exact 32-bit fixed point and use exp_nonpos
// (amp_impl.cuh).
//
// F64 (template parameter D): the same structure on fp64 values -- ONE codeword per CTA (its fp64 z plane takes the room
// of the two int32 planes, a group buffer is [8][512] doubles), LDS.64 gathers through tables scheduled for half-warp
// pools (amp.cu, build_pair_tables(f64 = 1)), A beta in fp64 registers, the reference's exact-equality stop rule, no
// quantisation.  It differs from the order-preserving STRICT kernel by fp64 summation-order noise only.
#include "amp_impl.cuh"

#include <type_traits>

// experiment switches (tools/ab_build2.sh): defaults are the measured best
#ifndef P2_FOLD_UNROLL
#define P2_FOLD_UNROLL 2  // unroll factor of the fold's loop over the 16 bins of a lane (code size vs scheduling freedom)
#endif
#ifndef P2_TW_UNROLL
#define P2_TW_UNROLL 1  // 1: one copy of the transform code for both slots; 2: one copy per slot
#endif
#ifndef P2_GATHER_DB
#define P2_GATHER_DB 1  // gather table words double-buffered in batches of 3 rows (0: batches of 6, no overlap)
#endif
#ifndef P2_SIGN_IMAD
#define P2_SIGN_IMAD 0  // gather: apply the sign with IMAD (fma pipe) instead of XOR + IADD3 (alu pipe)
#endif
#ifndef P2_PRELOAD
#define P2_PRELOAD 1  // first gather batch loaded before the CTA barrier that precedes the gather
#endif
#ifndef P2_REGS_TW
#define P2_REGS_TW 112  // setmaxnreg: registers per transform-warp thread (0 = leave 128/128); operator warps get 256 - this
#endif
#ifndef P2_PREFETCH_BETA
#define P2_PREFETCH_BETA 0  // prefetch.global.L2 of the next group's beta
#endif

namespace sb {
namespace p2 {

#ifdef P2_TRACE
// experiment builds (tools/ab_build2.sh TAG -DP2_TRACE): thread 0 of every CTA logs (pass start, pass end, end of the
// following iteration boundary) in ns and the slots' states; read with sb_p2_trace_read
constexpr int TRACE_MAX = 1 << 20;
__device__ unsigned long long g_trace[TRACE_MAX * 4];
__device__ int g_trace_n;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

constexpr int FOLD_UNROLL = P2_FOLD_UNROLL, TW_UNROLL = P2_TW_UNROLL;
constexpr int M = 512, S = 8;             // section size, sections per group
constexpr int NOW = 256, NR = 18;         // operator threads, rows of A beta per operator thread
constexpr int MAXN = NOW * NR;            // n <= 4608
constexpr int ZW = MAXN + 32;             // words per z plane: [n values | 32 zero words (one per bank) | unused]
constexpr int ZPLANE = ZW * 4;            // bytes
constexpr int SLOT = 4096;                // bytes per section of a group buffer: [2 codewords][512 int32]
constexpr int CWOFF = 2048;               // codeword 1 inside a section slot
constexpr int BUF = S * SLOT;             // 32 KB per group buffer
constexpr int SCR = 4096;                 // transpose scratch per transform warp

struct Slot {  // per-codeword state, shared memory; written by thread 0 between CTA barriers
    double tau2, zunit, fscale, funit, last_tau;
    int b, t, mode, first_zero, active, executed;  // mode 0: AMP iteration, 1: prologue z = y - A beta0
    unsigned flags;
    int pad;
};

struct Args {
    const uint16_t *inv2, *fwd2;
    const double *y, *Pl, *beta0;
    double *beta, *tau2_trace, *zscratch;
    int *iters, *n_exec;
    unsigned *flags;
    int *counter;  // work queue head (zeroed by the host before the launch) or NULL: codeword blockIdx.x only
    int L, n, T, B;
};

__host__ __device__ inline size_t smem_bytes(int L) {
    return 2 * (size_t)ZPLANE + 2 * (size_t)BUF + 8 * (size_t)SCR + 64 * sizeof(double) + 2 * sizeof(Slot) + 64 +
           sizeof(double) * (size_t)L;
}

__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
#ifdef P2_NO_BAR_OW  // timing experiment only (racy): upper bound of what a third group buffer could win
__device__ __forceinline__ void bar_ow() {}
#else
__device__ __forceinline__ void bar_ow() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
#endif

// deterministic CTA-wide reductions over all 16 warps (every thread adds the warp partials in warp order)
__device__ __forceinline__ double bsum(double v, double *scratch) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    bar_all();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    bar_all();
    double s = scratch[0];
#pragma unroll
    for (int i = 1; i < 16; i++) s += scratch[i];
    return s;
}
__device__ __forceinline__ double bmax(double v, double *scratch) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    bar_all();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    bar_all();
    double s = scratch[0];
#pragma unroll
    for (int i = 1; i < 16; i++) s = fmax(s, scratch[i]);
    return s;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ int lds32(const char *p) { return *reinterpret_cast<const int *>(p); }

__device__ __forceinline__ double ldsd(const char *p) { return *reinterpret_cast<const double *>(p); }

// 8 table entries (16 bytes) of one bin half: sums of the two codewords' z words they address
// (NC = 1: only the plane z0 points at -- the CTA's other slot is empty; D: fp64 plane, entries are k * 8)
template <int NC, bool D, typename V>
__device__ __forceinline__ void fold8(const uint4 w, const char *z0, V &a0, V &a1) {
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t lo = wd[i] & 0xFFFFu, hi = wd[i] >> 16;
        if constexpr (D) {
            a0 += ldsd(z0 + lo) + ldsd(z0 + hi);
        } else {
            a0 += lds32(z0 + lo) + lds32(z0 + hi);
            if (NC == 2) a1 += lds32(z0 + ZPLANE + lo) + lds32(z0 + ZPLANE + hi);
        }
    }
}

// operator warp `rw` folds z of both codewords into the 512 bins of section l and leaves them in the section's
// slot of the group buffer: word e*32+q of plane c = bin (q>>1)*32 + 2e + (q&1) (layout B of amp_impl.cuh)
template <int NC, bool D>
__device__ __forceinline__ void fold_section2(const uint16_t *__restrict__ inv2, int l, int q, const char *z0, char *slot) {
    using V = typename std::conditional<D, double, int>::type;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(inv2) + (size_t)l * (16 * 2 * 32) + q;
    V *st = reinterpret_cast<V *>(slot) + q;
    // The table words of the next bin are in flight while one bin is reduced.  The loop over the 16 bins is NOT
    // fully unrolled: the kernel's hot code has to fit the instruction caches (the SM's and the GPC's), see the
    // note on code size at role_main.
    uint4 c0 = __ldg(t4), c1 = __ldg(t4 + 32);
#pragma unroll FOLD_UNROLL
    for (int e = 0; e < 16; e++) {
        const int en = (e + 1 < 16) ? e + 1 : e;  // (the last iteration reloads its own words: no branch)
        const uint4 n0 = __ldg(t4 + (2 * en) * 32), n1 = __ldg(t4 + (2 * en + 1) * 32);
        V p0 = 0, p1 = 0, m0 = 0, m1 = 0;
        fold8<NC, D, V>(c0, z0, p0, p1);  // blocks of even parity: +
        fold8<NC, D, V>(c1, z0, m0, m1);  // blocks of odd parity: -
        st[e * 32] = p0 - m0;
        if (!D && NC == 2) st[CWOFF / 4 + e * 32] = p1 - m1;
        c0 = n0;
        c1 = n1;
    }
}

// prmt with a sign-replicating selector: every result byte = the msb of byte 1 (0x9999) or byte 3 (0xBBBB) of w
template <unsigned SEL>
__device__ __forceinline__ int sign_mask(uint32_t w) {
    int m;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(w), "r"(SEL));
    return m;
}

// 8 entries of (group, row): signed words of both codewords' planes (D: fp64 words, sign = bit 15 -> bit 63)
template <int NC, bool D, typename V>
__device__ __forceinline__ void gather8(const uint4 w, const char *bufg, V &p0, V &p1) {
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if constexpr (D) {
            const double vlo = ldsd(bufg + (wd[i] & 0x7FF8u)), vhi = ldsd(bufg + ((wd[i] >> 16) & 0x7FF8u));
            p0 += __hiloint2double(__double2hiint(vlo) ^ (int)((wd[i] << 16) & 0x80000000u), __double2loint(vlo)) +
                  __hiloint2double(__double2hiint(vhi) ^ (int)(wd[i] & 0x80000000u), __double2loint(vhi));
        } else {
            const uint32_t olo = wd[i] & 0x7FFCu, ohi = (wd[i] >> 16) & 0x7FFCu;
#if P2_SIGN_IMAD
            // +-1 from bit 15 / bit 31; the signed add is one IMAD on the (idle) fma pipe instead of XOR + IADD3 on the alu pipe
            const int slo = sign_mask<0x9999>(wd[i]) | 1, shi = sign_mask<0xBBBB>(wd[i]) | 1;
            p0 += lds32(bufg + olo) * slo + lds32(bufg + ohi) * shi;
            if (NC == 2) p1 += lds32(bufg + CWOFF + olo) * slo + lds32(bufg + CWOFF + ohi) * shi;
#else
            const int mlo = sign_mask<0x9999>(wd[i]), mhi = sign_mask<0xBBBB>(wd[i]);  // 0 / -1 from bit 15 / bit 31
            p0 += ((lds32(bufg + olo) ^ mlo) - mlo) + ((lds32(bufg + ohi) ^ mhi) - mhi);
            if (NC == 2) p1 += ((lds32(bufg + CWOFF + olo) ^ mlo) - mlo) + ((lds32(bufg + CWOFF + ohi) ^ mhi) - mhi);
#endif
        }
    }
}

constexpr int KB = P2_GATHER_DB ? 3 : 6;  // rows per gather batch

// table words of the first row batch of group g for operator thread `ot` (issued a barrier ahead of their use)
__device__ __forceinline__ void gather_preload(const uint16_t *__restrict__ fwd2, int g, int n, int ot, uint4 (&w)[KB]) {
    const uint4 *tab = reinterpret_cast<const uint4 *>(fwd2) + (size_t)g * n + ot;
#pragma unroll
    for (int j = 0; j < KB; j++) w[j] = (ot + NOW * j < n) ? __ldg(tab + NOW * j) : make_uint4(0, 0, 0, 0);
}

// operator thread `ot` adds the 8 sections of group g into its rows k = ot + 256 j of both codewords; the table
// words of batch b+1 are in flight while batch b is gathered
template <int NC, bool D, typename A>
__device__ __forceinline__ void gather_group2(const uint16_t *__restrict__ fwd2, int g, int n, int ot, const char *bufg,
                                              A (&acc)[NR][2], const uint4 (&w0)[KB]) {
    using V = typename std::conditional<D, double, int>::type;
    const uint4 *tab = reinterpret_cast<const uint4 *>(fwd2) + (size_t)g * n + ot;
    uint4 w[2][KB];
#pragma unroll
    for (int j = 0; j < KB; j++) w[0][j] = w0[j];
#pragma unroll
    for (int b = 0; b < NR / KB; b++) {
#if !P2_GATHER_DB
        if (b > 0) {
#pragma unroll
            for (int j = 0; j < KB; j++) {
                const int r = b * KB + j;
                w[b & 1][j] = (ot + NOW * r < n) ? __ldg(tab + NOW * r) : make_uint4(0, 0, 0, 0);
            }
        }
#else
        if (b + 1 < NR / KB) {
#pragma unroll
            for (int j = 0; j < KB; j++) {
                const int r = (b + 1) * KB + j;
                w[(b + 1) & 1][j] = (ot + NOW * r < n) ? __ldg(tab + NOW * r) : make_uint4(0, 0, 0, 0);
            }
        }
#endif
#pragma unroll
        for (int j = 0; j < KB; j++) {
            V p0 = 0, p1 = 0;  // 8 terms (FAST: of < 2^27 each)
            gather8<NC, D, V>(w[b & 1][j], bufg, p0, p1);
            acc[b * KB + j][0] += p0;
            if (!D && NC == 2) acc[b * KB + j][1] += p1;
        }
    }
}

// transform warp: one section of one codeword.  mode 0: fold result -> FHT -> softmax -> beta -> FHT -> F;
// mode 1: beta0 -> (copy to beta) -> FHT -> F.  Same arithmetic and order as section_phase<9, *, true> (TRQ path).
template <bool D>
__device__ __forceinline__ void transform_unit(const Args &a, const Slot *sl, int sidx, int q, void *stv, int *Sp, int *Sn,
                                               const double *rtp, double inv_rt_n, double &sq, double &gmax, double &lmin) {
    double x[16];
    int *st = static_cast<int *>(stv);
    double *std_ = static_cast<double *>(stv);  // D: the slot holds 512 doubles (fold result in, F out)
    const size_t boff = ((size_t)sl->b * a.L + sidx) * M;
    if (sl->mode == 0) {
        const double zunit = sl->zunit;
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = D ? std_[e * 32 + q] : (double)st[e * 32 + q] * zunit;
        double bv[16];
        const bool fz = sl->first_zero != 0;
        const double *bsrc = a.beta + boff;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            bv[e] = 0.0;
            if (!fz) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(bv[e]) : "l"(bsrc + e * 32 + q));
        }
        fht512_B_to_A(x, q, Sp, Sn);
        const double c2 = rtp[sidx] / sl->tau2;
        double m = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const double s = bv[e] + x[e] * inv_rt_n;  // s = beta + A^T z          (sparc_ldpc.py:213)
            x[e] = s * c2;                            // u = s sqrt(n P_l)/tau^2    (:215)
            m = fmax(m, x[e]);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
        gmax = fmax(gmax, m);
        lmin = fmin(lmin, m);
        double sum = 0.0;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            // section max instead of the reference's global max (:216): same softmax
            x[e] = (!D && SB_FAST_EXP) ? exp_nonpos(x[e] - m) : exp(x[e] - m);
            sum += x[e];
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const double sc = rtp[sidx] / sum;
        double *bdst = a.beta + boff;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            x[e] *= sc;  // beta = sqrt(n P_l) softmax(u)    (:218-219)
            sq += x[e] * x[e];
            __stcs(bdst + e * 32 + q, x[e]);
        }
    } else {
        const double *bsrc = a.beta0 + boff;
        double *bdst = a.beta + boff;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            x[e] = bsrc[e * 32 + q];
            bdst[e * 32 + q] = x[e];
        }
    }
    if constexpr (D) {
        fht512_A_to_B(x, q, Sp, Sn);
#pragma unroll
        for (int e = 0; e < 16; e++) std_[e * 32 + q] = x[e];  // word e*32+q = fq_word(lo)
    } else {
#if SB_INT_FHT2  // exact 32-bit fixed-point transform of the 30-bit quantised beta (amp_impl.cuh), F rounded to 27 bits
        int xi[16];
        const double fs30 = sl->fscale * 8.0;
#pragma unroll
        for (int e = 0; e < 16; e++) xi[e] = __double2int_rn(x[e] * fs30);
        fht512_A_to_B_i(xi, q, Sp);
#pragma unroll
        for (int e = 0; e < 16; e++) st[e * 32 + q] = (xi[e] + 4) >> 3;  // word e*32+q = fq_word(lo)
#else
        fht512_A_to_B(x, q, Sp, Sn);
        const double fs = sl->fscale;
#pragma unroll
        for (int e = 0; e < 16; e++) st[e * 32 + q] = __double2int_rn(x[e] * fs);  // word e*32+q = fq_word(lo)
#endif
    }
}

// Operator warps, one pass over the sections: fold group i+1 and gather group i-1 while the transform warps work on
// group i (same barrier sequence as the transform warps' loop in role_main).  NC = 2: both slots, planes at z0 /
// z0 + ZPLANE and buf / buf + CWOFF; NC = 1: one slot, z0 and buf point at its planes, sums in acc[.][0].
template <int NC, bool D, typename A>
__device__ __forceinline__ void ow_pass(const Args &a, int n, int G, int rt, int rw, int q, const char *z0, char *buf,
                                        bool anyfold, A (&acc)[NR][2]) {
    uint4 pre[KB];  // first table words of the next gather
#pragma unroll
    for (int j = 0; j < NR; j++) acc[j][0] = acc[j][1] = 0;
    if (anyfold) fold_section2<NC, D>(a.inv2, rw, q, z0, buf + rw * SLOT);
    if (P2_PRELOAD) gather_preload(a.fwd2, 0, n, rt, pre);
    bar_all();
    for (int i = 0;; i++) {  // (one copy of the gather code: the last trip gathers group G-1 and leaves)
        if (i >= 1) {
            if (!P2_PRELOAD) gather_preload(a.fwd2, i - 1, n, rt, pre);
            gather_group2<NC, D, A>(a.fwd2, i - 1, n, rt, buf + ((i - 1) & 1) * BUF, acc, pre);
        }
        if (i == G) break;
        bar_ow();  // every operator warp has finished reading that buffer before group i+1 is folded into it
        if (i + 1 < G && anyfold) fold_section2<NC, D>(a.inv2, (i + 1) * S + rw, q, z0, buf + ((i + 1) & 1) * BUF + rw * SLOT);
        if (P2_PRELOAD && i >= 1) gather_preload(a.fwd2, i, n, rt, pre);
        bar_all();
    }
}

template <bool OW, bool D>
__device__ __forceinline__ void role_main(const Args &a, unsigned char *smem, const double P, const double fscale_q) {
    constexpr int NSLOT = D ? 1 : 2;  // D: one fp64 codeword per CTA (its z plane takes the room of the two int32 planes)
    using A = typename std::conditional<D, double, long long>::type;
    const int tid = threadIdx.x, q = tid & 31, n = a.n, L = a.L, G = L / S;
    const int rt = OW ? tid : tid - NOW, rw = rt >> 5;  // thread / warp index inside the role
    char *zq = reinterpret_cast<char *>(smem);
    char *buf = zq + 2 * ZPLANE;
    char *scr = buf + 2 * BUF;
    double *red = reinterpret_cast<double *>(scr + 8 * SCR);
    Slot *slot = reinterpret_cast<Slot *>(red + 64);
    int *misc = reinterpret_cast<int *>(slot + 2);
    double *rtp = reinterpret_cast<double *>(misc + 16);
    const double nd = (double)n, rt_n = sqrt(nd), inv_rt_n = 1.0 / rt_n;
    // fp64 z of the slots.  FAST: [2][n] in global scratch, rows owned by one thread (the shared planes hold the
    // fixed-point copies); D: the shared plane itself
    double *zf = D ? reinterpret_cast<double *>(zq) : a.zscratch + (size_t)blockIdx.x * 2 * n;
    A acc[NR][2];  // operator warps only (D: [.][0] only)
    bool exhausted = false;

    for (;;) {
        // ---------------- iteration boundary: refill free slots, tau and the stop rule, fixed-point copy of z
#pragma unroll 1
        for (int c = 0; c < NSLOT; c++) {
            Slot *sl = slot + c;
            double *zc = zf + (size_t)c * n;
            for (;;) {
                if (!sl->active) {
                    if (exhausted) break;
                    if (tid == 0) {
                        int b;
                        if (a.counter) b = atomicAdd(a.counter, 1);
                        else b = (misc[1]++ == 0) ? (int)blockIdx.x : a.B;
                        misc[0] = b;
                    }
                    bar_all();
                    const int b = misc[0];
                    bar_all();
                    if (b >= a.B) { exhausted = true; break; }
                    if (tid == 0) {
                        sl->tau2 = 1.0; sl->zunit = 1.0; sl->fscale = fscale_q; sl->funit = 1.0 / fscale_q; sl->last_tau = 0.0;
                        sl->b = b; sl->t = 0; sl->mode = a.beta0 ? 1 : 0; sl->first_zero = a.beta0 ? 0 : 1;
                        sl->active = 1; sl->executed = 0; sl->flags = 0;
                    }
                    if (a.beta0 == nullptr) {
                        if (OW) {
                            const double *y = a.y + (size_t)b * n;
                            for (int k = rt; k < n; k += NOW) zc[k] = y[k];
                        }
                        bar_all();
                    } else {
                        if constexpr (!D) {  // |FHT_M(beta0_l)| <= sum_j |beta0_l[j]|: one streaming pass gives the fixed-point scale
                            const double *b0 = a.beta0 + (size_t)b * L * M;
                            double bound = 0.0;
                            for (int sidx = tid >> 5; sidx < L; sidx += 16) {
                                double s1 = 0.0;
                                for (int j = q; j < M; j += 32) s1 += fabs(b0[(size_t)sidx * M + j]);
#pragma unroll
                                for (int d = 16; d >= 1; d >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, d);
                                bound = fmax(bound, s1);
                            }
                            const double bm = bmax(bound, red);
                            if (tid == 0) {
                                const double fs = scalbn(1.0, 27 - ceil_exp(bm * (1.0 + 1e-6)));
                                sl->fscale = fs;
                                sl->funit = 1.0 / fs;
                            }
                        }
                        bar_all();  // the slot's fields (thread 0) are visible to everyone
                    }
                }
                if (sl->mode == 1) break;  // the prologue pass comes first
                bool done = sl->t >= a.T;
                double tau = 0.0;
                if (!done) {
                    // the thread's rows of z, all loads in flight together (FAST: z lives in global scratch; a rolled loop
                    // exposed one L2 round trip per row, twice per slot: 19 -> 12 us per iteration boundary) and kept
                    // for the fixed-point copy below (the accumulator registers are dead here)
                    double part = 0.0, zmax = 0.0, zv[NR];
                    if (OW) {
#pragma unroll
                        for (int j = 0; j < NR; j++) {
                            const int k = rt + NOW * j;
                            zv[j] = (k < n) ? zc[k] : 0.0;
                        }
#pragma unroll
                        for (int j = 0; j < NR; j++) {
                            part += zv[j] * zv[j];
                            zmax = fmax(zmax, fabs(zv[j]));
                        }
                    }
                    tau = sqrt(bsum(part, red) / nd);  // (:203)
                    const double lt = sl->last_tau;
                    if (D) bar_all();  // everyone has read last_tau before thread 0 overwrites it (FAST: bmax below)
                    // FAST stop rule of amp_kernel: tau jitters at the quantisation floor instead of reaching an exact
                    // fp64 fixed point (:204), so stop once it moves by less than 2^-27 relative; D: the reference's
                    // exact-equality rule
                    if (tau == lt || (!D && fabs(tau - lt) <= tau * 7.450580596923828e-09)) {
                        done = true;
                        if (tid == 0) sl->flags |= SB_AMP_STOPPED;
                    } else if (D) {
                        if (tid == 0) {
                            sl->last_tau = tau;
                            sl->tau2 = tau * tau;
                            if (a.tau2_trace != nullptr) a.tau2_trace[(size_t)sl->b * a.T + sl->t] = tau * tau;
                        }
                    } else {
                        const int ez = ceil_exp(bmax(zmax, red) * (1.0 + 1e-6));
                        const double zscale = scalbn(1.0, 27 - ez);
                        if (tid == 0) {
                            sl->last_tau = tau;
                            sl->tau2 = tau * tau;
                            sl->zunit = scalbn(1.0, ez - 27);
                            if (a.tau2_trace != nullptr) a.tau2_trace[(size_t)sl->b * a.T + sl->t] = tau * tau;
                        }
                        if (OW) {
                            int *zp = reinterpret_cast<int *>(zq + c * ZPLANE);
#pragma unroll
                            for (int j = 0; j < NR; j++) {
                                const int k = rt + NOW * j;
                                if (k < n) zp[k] = __double2int_rn(zv[j] * zscale);
                            }
                        }
                    }
                }
                if (!done) break;
                // ---- codeword finished: results out, slot free
                {
                    const int b = sl->b;
                    if (sl->first_zero) {  // T == 0 or stop before the first update: beta is the zero vector
                        double *beta = a.beta + (size_t)b * L * M;
                        for (int i = tid; i < L * M; i += 512) beta[i] = 0.0;
                    }
                    bar_all();
                    if (tid == 0) {
                        const int t = sl->t;
                        a.iters[b] = (t < a.T) ? t : (a.T > 0 ? a.T - 1 : 0);
                        a.n_exec[b] = sl->executed;
                        a.flags[b] = sl->flags;
                        sl->active = 0;
                    }
                    bar_all();
                }
            }
        }
        bar_all();
        const int m0 = slot[0].active ? slot[0].mode : -1, m1 = (!D && slot[1].active) ? slot[1].mode : -1;
        if (m0 < 0 && m1 < 0) break;
        const bool anyfold = (m0 == 0) || (m1 == 0);

        // ---------------- one pass over the sections, software-pipelined by role
        double sq0 = 0.0, sq1 = 0.0, gmax0 = -INFINITY, gmax1 = -INFINITY, lmin0 = INFINITY, lmin1 = INFINITY;
        const bool both = (m0 >= 0) && (m1 >= 0);
#ifdef P2_TRACE
        int trow = -1;
        if (tid == 0) {
            trow = atomicAdd(&g_trace_n, 1);
            if (trow < TRACE_MAX) {
                g_trace[trow * 4 + 0] = gtime();
                g_trace[trow * 4 + 3] = ((unsigned long long)blockIdx.x << 32) | ((unsigned)(m0 + 1) << 28) | ((unsigned)(m1 + 1) << 24) |
                                        ((unsigned)(m0 >= 0 ? slot[0].t : 0) << 8) | (unsigned)(m1 >= 0 ? slot[1].t : 0);
            }
        }
#endif
        if constexpr (OW) {
            if constexpr (D) ow_pass<1, true, A>(a, n, G, rt, rw, q, zq, buf, anyfold, acc);
            else if (both) ow_pass<2, false, A>(a, n, G, rt, rw, q, zq, buf, anyfold, acc);
            else ow_pass<1, false, A>(a, n, G, rt, rw, q, zq + (m0 >= 0 ? 0 : ZPLANE), buf + (m0 >= 0 ? 0 : CWOFF), anyfold, acc);
        } else {
            bar_all();
            for (int i = 0; i < G; i++) {
                char *sec = buf + (i & 1) * BUF + rw * SLOT;
                int *Sp = reinterpret_cast<int *>(scr + rw * SCR), *Sn = Sp + SCR / 8;
                if (P2_PREFETCH_BETA && i + 1 < G) {  // beta of the next group: HBM -> L2 while this group is transformed
                    const size_t nxt = ((size_t)((i + 1) * S + rw)) * M + q * 16;
                    if (m0 == 0 && !slot[0].first_zero) prefetch_l2(a.beta + (size_t)slot[0].b * L * M + nxt);
                    if (m1 == 0 && !slot[1].first_zero) prefetch_l2(a.beta + (size_t)slot[1].b * L * M + nxt);
                }
#pragma unroll TW_UNROLL
                for (int c = 0; c < NSLOT; c++) {  // (one copy of the transform code: instruction-cache footprint)
                    if ((c ? m1 : m0) < 0) continue;
                    double dsq = 0.0, dmax = -INFINITY, dmin = INFINITY;
                    transform_unit<D>(a, slot + c, i * S + rw, q, sec + c * CWOFF, Sp, Sn, rtp, inv_rt_n, dsq, dmax, dmin);
                    if (c) { sq1 += dsq; gmax1 = fmax(gmax1, dmax); lmin1 = fmin(lmin1, dmin); }
                    else { sq0 += dsq; gmax0 = fmax(gmax0, dmax); lmin0 = fmin(lmin0, dmin); }
                }
                bar_all();
            }
        }

#ifdef P2_TRACE
        if (tid == 0 && trow >= 0 && trow < TRACE_MAX) g_trace[trow * 4 + 1] = gtime();
#endif
        // ---------------- end of pass: Onsager term and residual (sparc_ldpc.py:220), or z = y - A beta0 (:197-198)
#pragma unroll
        for (int c = 0; c < NSLOT; c++) {
            const int mc = c ? m1 : m0;
            if (mc < 0) continue;
            Slot *sl = slot + c;
            double *zc = zf + (size_t)c * n;
            const double *y = a.y + (size_t)sl->b * n;
            const double funit = sl->funit;
            if (mc == 1) {
                if constexpr (OW) {
#pragma unroll
                    for (int j = 0; j < NR; j++) {
                        const int k = rt + NOW * j;
                        if (k < n) zc[k] = y[k] - ((double)((both && c) ? acc[j][1] : acc[j][0]) * funit) / rt_n;
                    }
                }
                bar_all();
                if (tid == 0) {
                    sl->mode = 0;
                    sl->fscale = fscale_q;
                    sl->funit = 1.0 / fscale_q;
                }
            } else {
                const double sumsq = bsum(c ? sq1 : sq0, red);
                const double gm = bmax(c ? gmax1 : gmax0, red);
                const double lm = -bmax(-(c ? lmin1 : lmin0), red);
                const double ons = P - sumsq / nd, tau2 = sl->tau2;
                if constexpr (OW) {
#pragma unroll
                    for (int j = 0; j < NR; j++) {
                        const int k = rt + NOW * j;
                        if (k < n) zc[k] = (y[k] - ((double)((both && c) ? acc[j][1] : acc[j][0]) * funit) / rt_n) + (zc[k] / tau2) * ons;
                    }
                }
                bar_all();
                if (tid == 0) {
                    // the reference subtracts the GLOBAL maximum (:216): see SB_AMP_REF_NAN in sparc_b200.h
                    if (gm - lm > 708.39) sl->flags |= SB_AMP_REF_NAN;
                    sl->executed++;
                    sl->t++;
                    sl->first_zero = 0;
                }
            }
        }
        bar_all();
#ifdef P2_TRACE
        if (tid == 0 && trow >= 0 && trow < TRACE_MAX) g_trace[trow * 4 + 2] = gtime();
#endif
    }
}

template <bool D>
__global__ void __launch_bounds__(512, 1) amp2_kernel(Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, n = a.n, L = a.L;
    char *zq = reinterpret_cast<char *>(smem);
    double *red = reinterpret_cast<double *>(zq + 2 * ZPLANE + 2 * BUF + 8 * SCR);
    Slot *slot = reinterpret_cast<Slot *>(red + 64);
    int *misc = reinterpret_cast<int *>(slot + 2);
    double *rtp = reinterpret_cast<double *>(misc + 16);
    if (D) {
        if (tid < 32) reinterpret_cast<double *>(zq)[n + tid] = 0.0;  // 32 zero words behind the fp64 plane
    } else if (tid < 64) {
        reinterpret_cast<int *>(zq + (tid >> 5) * ZPLANE)[n + (tid & 31)] = 0;  // one zero word per bank and plane
    }
    if (tid < 2) {
        slot[tid].active = 0;
        slot[tid].mode = 0;
    }
    if (tid == 0) misc[1] = 0;
    const double nd = (double)n;
    double pl = 0.0, plmax = 0.0;
    for (int i = tid; i < L; i += 512) {
        const double p = a.Pl[i];
        pl += p;
        plmax = fmax(plmax, p);
        rtp[i] = sqrt(nd * p);  // sqrt(n P_l), (:215)
    }
    __syncthreads();
    const double P = bsum(pl, red);  // np.sum(Pl) (:190)
    const double cmax = sqrt(nd * bmax(plmax, red));
    // |F| <= sqrt(n P_l) <= cmax; the 1e-6 margin keeps |F_q| strictly below 2^27 (sums of 8 terms stay in int32)
    const double fscale_q = D ? 1.0 : scalbn(1.0, 27 - ceil_exp(cmax * (1.0 + 1e-6)));
    bar_all();
#ifdef P2_DESYNC  // experiment: start the CTAs out of phase
    for (int i = 0; i < (int)(blockIdx.x % 37); i++) __nanosleep(20000);
    __syncthreads();
#endif
    if (tid < NOW) {
#if P2_REGS_TW
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(256 - P2_REGS_TW));
#endif
        role_main<true, D>(a, smem, P, fscale_q);
    } else {
#if P2_REGS_TW
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(P2_REGS_TW));
#endif
        role_main<false, D>(a, smem, P, fscale_q);
    }
}

}  // namespace p2

#ifdef P2_TRACE
extern "C" int sb_p2_trace_read(unsigned long long *out, int max_rows) {
    int n = 0;
    SB_CUDA(cudaDeviceSynchronize());
    SB_CUDA(cudaMemcpyFromSymbol(&n, p2::g_trace_n, sizeof(int)));
    if (n > p2::TRACE_MAX) n = p2::TRACE_MAX;
    if (n > max_rows) n = max_rows;
    SB_CUDA(cudaMemcpyFromSymbol(out, p2::g_trace, sizeof(unsigned long long) * 4 * (size_t)n));
    const int zero = 0;
    SB_CUDA(cudaMemcpyToSymbol(p2::g_trace_n, &zero, sizeof(int)));
    return n;
}
#endif

// FAST mode, all sections active, pair tables present.  scratch: [2][B][n] doubles (sb_amp_batch): the first
// 2 * grid * n hold z of the resident slots, the work counter sits behind them when B > grid.
int launch_amp2(const sb_operator *op, const AmpArgs &aa, int B, int f64, cudaStream_t st) {
    static int nsm = 0;
    if (nsm == 0) {
        int dev = 0, v = 0;
        SB_CUDA(cudaGetDevice(&dev));
        SB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        nsm = v > 0 ? v : 148;
    }
    p2::Args a;
    memset(&a, 0, sizeof(a));
    a.inv2 = f64 ? op->inv2d : op->inv2; a.fwd2 = f64 ? op->fwd2d : op->fwd2;
    a.y = aa.y; a.Pl = aa.Pl; a.beta0 = aa.beta0; a.beta = aa.beta; a.tau2_trace = aa.tau2_trace; a.zscratch = aa.zscratch;
    a.iters = aa.iters; a.n_exec = aa.n_exec; a.flags = aa.flags;
    a.L = op->L; a.n = op->n; a.T = aa.T; a.B = B;
    const int grid = B < nsm ? B : nsm;
    if (B > grid) {
        a.counter = reinterpret_cast<int *>(aa.zscratch + (size_t)2 * grid * op->n);
        SB_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(int), st));
    }
    const size_t smem = p2::smem_bytes(op->L);
    if (smem > 227 * 1024) return fail(SB_EINVAL, "AMP (pair kernel): L too large for shared memory%s (%ld bytes)", "", (long)smem);
    if (f64) {
        SB_CUDA(cudaFuncSetAttribute(p2::amp2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        p2::amp2_kernel<true><<<grid, 512, smem, st>>>(a);
    } else {
        SB_CUDA(cudaFuncSetAttribute(p2::amp2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        p2::amp2_kernel<false><<<grid, 512, smem, st>>>(a);
    }
    SB_LAUNCHED();
    return SB_OK;
}

}  // namespace sb
