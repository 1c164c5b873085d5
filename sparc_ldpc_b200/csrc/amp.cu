// amp.cu -- host side of the SPARC design operator / AMP entry points: table construction and dispatch.
// The kernels live in amp_impl.cuh and are instantiated per section size in amp_inst_*.cu.
#include "amp_impl.cuh"
#include "sched.h"

#include <functional>
#include <mutex>
#include <thread>
#include <vector>

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

// ---- FAST-mode tables: conflict-free orderings of the two shared-memory gathers (sched.h) -------------------
// Host-only; sb_operator_create uploads the result, sb_fast_tables_check (tests) verifies it without a GPU.
struct FastTables {
    int qok = 0, qpre = 0, qneg = 0, PW = 0, GQ = 0;  // PW = sections per gather pool (8 | 16), GQ = floor(L / PW)
    std::vector<uint16_t> invq, fwdq;
    long fold_steps = 0, fold_wavefronts = 0, gather_steps = 0, gather_wavefronts = 0;  // bank-conflict model
};

static int team_lanes(int M) { return M >= 128 ? 32 : (M >= 4 ? M / 4 : 1); }

// bin held by register e of lane q after the fold: M = 512 uses the transposed layout B of amp_impl.cuh
static int fast_bin(int logM, int TEAM, int e, int q) {
    return logM == 9 ? (q >> 1) * 32 + 2 * e + (q & 1) : e * TEAM + q;
}

static void parallel_for(int count, const std::function<void(int, int)> &body) {
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    if (nt > count) nt = count;
    if (nt <= 1) { body(0, count); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const int a = (int)((long)count * t / nt), b = (int)((long)count * (t + 1) / nt);
        th.emplace_back(body, a, b);
    }
    for (auto &x : th) x.join();
}

// shared memory of the FAST kernel with W section slots (= Smem<LOGM, true>::bytes, logM <= 9)
static size_t fast_smem_bytes(int logM, int n, int W) {
    const size_t region = (size_t)2 << sign_bit(logM);
    return 4 * (size_t)pad4(2 * n + 32) + 8 * (size_t)pad2(n) + (size_t)((W + 7) / 8) * region + 8 * 40 + 4 * (size_t)(W + 2);
}

// Table layouts (16-bit entries, every lane reads 16 bytes = 8 entries per load and the 32 lanes of a warp read
// 512 contiguous bytes):
//   invq [L][EPT][Hp/8][TEAM][8]   entry t = h*8+i of bin e*TEAM+q: offset into [ +z | 32 zero words | -z ]
//   fwdq [GQ][PW/8][n][8]          entry t = h*8+i of (chunk, row k): byte offset into the chunk's +-F area
static void build_fast_tables(const uint32_t *ordering, int L, int M, int n, int Hp, FastTables &ft) {
    const int logM = ilog2(M), SBQ = sign_bit(logM), TEAM = team_lanes(M), EPT = M / TEAM, NH = Hp / 8;
    ft.qneg = n + 32;
    if (2L * n + 32 > 65535) return;  // offsets do not fit 16 bits: FAST falls back to STRICT
    ft.qok = 1;
    ft.qpre = ((2L * n + 32) * 4 <= 65535) ? 1 : 0;
    const int osh = ft.qpre ? 2 : 0;
    ft.invq.assign((size_t)L * M * Hp, 0);
    std::vector<long> stat((size_t)L * 2, 0);
    // fold: pool = (section, e): lane q of the team owns bin e*TEAM + q and reads one term per step
    parallel_for(L, [&](int l0, int l1) {
        PoolScheduler ps;
        std::vector<PoolEdge> E;
        std::vector<std::vector<int>> bins(M);
        std::vector<int> used((size_t)Hp * 32);
        for (int l = l0; l < l1; l++) {
            for (auto &b : bins) b.clear();
            for (int k = 0; k < n; k++) {
                const uint32_t r = ordering[(size_t)l * n + k];
                const int neg = __builtin_popcount(r / M) & 1;  // sign of block hi in the fold tree = (-1)^popcount(hi)
                bins[r % M].push_back(neg ? ft.qneg + k : k);
            }
            for (int e = 0; e < EPT; e++) {
                E.clear();
                for (int q = 0; q < TEAM; q++)
                    for (int o : bins[fast_bin(logM, TEAM, e, q)]) E.push_back(PoolEdge{q, o & 31, o, 0});
                ps.run(E, Hp);
                PoolScheduler::improve(E, Hp);
                stat[(size_t)l * 2] += Hp;
                stat[(size_t)l * 2 + 1] += PoolScheduler::cost(E, Hp);
                std::fill(used.begin(), used.end(), 0);
                for (const PoolEdge &pe : E) used[(size_t)pe.step * 32 + pe.bank] = 1;
                // idle slots read a zero word that lies in a bank no real term of the step uses
                for (int t = 0; t < Hp; t++) {
                    int fb = 0;
                    for (int b = 0; b < 32; b++) if (!used[(size_t)t * 32 + b]) { fb = b; break; }
                    const int zero_word = n + ((fb - n) & 31);  // word in [n, n+32) with bank fb
                    for (int q = 0; q < TEAM; q++)
                        ft.invq[((((size_t)l * EPT + e) * NH + (t >> 3)) * TEAM + q) * 8 + (t & 7)] = (uint16_t)(zero_word << osh);
                }
                for (const PoolEdge &pe : E)
                    ft.invq[((((size_t)l * EPT + e) * NH + (pe.step >> 3)) * TEAM + pe.lane) * 8 + (pe.step & 7)] =
                        (uint16_t)(pe.id << osh);
            }
        }
    });
    for (int l = 0; l < L; l++) { ft.fold_steps += stat[(size_t)l * 2]; ft.fold_wavefronts += stat[(size_t)l * 2 + 1]; }
    // gather: pool = (chunk of PW sections, warp row of 32 consecutive k); lane k%32 reads one section per step.
    // PW = 16 (one 512-thread CTA per SM, 1.75 wavefronts per step) measured faster than PW = 8 (two 256-thread
    // CTAs per SM, 2.0 wavefronts per step): 3.49 vs 3.68 us per codeword-iteration at L=M=512.  SB_AMP_POOL=8
    // selects the latter for experiments when two CTAs fit.
    if (logM > 9) return;
    ft.PW = 16;
    if (const char *env = knob("SB_AMP_POOL"))
        if (atoi(env) == 8 && TEAM == 32 && 2 * (fast_smem_bytes(logM, n, 8) + 1024) <= 227 * 1024) ft.PW = 8;
    const int PW = ft.PW;
    ft.GQ = L / PW;
    if (ft.GQ == 0) return;
    ft.fwdq.assign((size_t)ft.GQ * n * PW, 0);
    const int SPR = 8;
    std::vector<long> gstat((size_t)ft.GQ * 2, 0);
    parallel_for(ft.GQ, [&](int g0, int g1) {
        PoolScheduler ps;
        std::vector<PoolEdge> E;
        for (int g = g0; g < g1; g++)
            for (int k0 = 0; k0 < n; k0 += 32) {
                E.clear();
                const int nl = (n - k0 < 32) ? n - k0 : 32;
                for (int q = 0; q < nl; q++)
                    for (int i = 0; i < PW; i++) {
                        const uint32_t r = ordering[(size_t)(g * PW + i) * n + k0 + q];
                        const uint32_t lo = r % M, sg = __builtin_popcount(r / M) & 1;
                        const uint32_t off = (uint32_t)(i / SPR) * (2u << SBQ) + (uint32_t)(i % SPR) * ((uint32_t)M << 2) +
                                             (fq_word(logM, lo) << 2) + (sg << SBQ);
                        E.push_back(PoolEdge{q, (int)((off >> 2) & 31), (int)off, 0});
                    }
                ps.run_best(E, PW);
                gstat[(size_t)g * 2] += PW;
                gstat[(size_t)g * 2 + 1] += PoolScheduler::cost(E, PW);
                for (const PoolEdge &pe : E)
                    ft.fwdq[(((size_t)g * (PW / 8) + (pe.step >> 3)) * n + k0 + pe.lane) * 8 + (pe.step & 7)] = (uint16_t)pe.id;
            }
    });
    for (int g = 0; g < ft.GQ; g++) { ft.gather_steps += gstat[(size_t)g * 2]; ft.gather_wavefronts += gstat[(size_t)g * 2 + 1]; }
}


// ---- pair-kernel tables (amp2.cu: two codewords per CTA, M = 512, w/M <= 16, n <= 4608, L % 8 == 0) ---------
// The same two maps as build_fast_tables with three differences: (1) there is no -z / -F copy: the fold's sign is
// the HALF of the bin's 16 steps an entry sits in (steps 0..7 are added, 8..15 subtracted; a bin has at most 8
// blocks of each parity when w/M <= 16), the gather's sign is bit 15 of the entry; (2) a gather chunk is 8
// sections: entry = slot * 4096 + fq_word(lo) * 4 | sign << 15, the byte offset of codeword 0's word inside the
// group buffer [8 slots][2 codewords][512 int32] (codeword 1 sits 2048 bytes further); (3) fold entries are byte
// offsets into one codeword's z plane [n values | 32 zero words].
struct PairTables {
    int ok = 0;
    std::vector<uint16_t> inv2, fwd2;
    long fold_steps = 0, fold_wavefronts = 0, gather_steps = 0, gather_wavefronts = 0;
};

static bool pair_shape_ok(int L, int M, int n, int H) { return M == 512 && H <= 16 && n <= 4608 && n >= 32 && L % 8 == 0; }

// f64 = 0: 4-byte words, pools of 32 lanes x 32 banks (the FAST pair kernel).
// f64 = 1: the same kernel structure on fp64 values (SB_AMP_F64: one codeword per CTA, 8-byte words): a warp-wide
//          LDS.64 is served half-warp by half-warp, so a pool is 16 lanes x 16 eight-byte banks; fold entries are
//          byte offsets k * 8 into the fp64 z plane [n | 32 zero words], gather entries slot * 4096 + word * 8 | sign
//          << 15 into a group buffer [8 slots][512 doubles].
static void build_pair_tables(const uint32_t *ordering, int L, int M, int n, int H, PairTables &pt, int f64 = 0) {
    if (!pair_shape_ok(L, M, n, H)) return;
    pt.ok = 1;
    const int logM = 9, TEAM = 32, EPT = 16;
    const int LG = f64 ? 16 : 32, ESZ = f64 ? 8 : 4;  // lanes (= banks) per pool, bytes per element
    pt.inv2.assign((size_t)L * M * 16, 0);
    std::vector<long> stat((size_t)L * 2, 0);
    parallel_for(L, [&](int l0, int l1) {
        PoolScheduler ps;
        std::vector<PoolEdge> E;
        std::vector<std::vector<int>> bins(2 * M);  // [sign][bin]
        std::vector<int> used(8 * 32);
        for (int l = l0; l < l1; l++) {
            for (auto &b : bins) b.clear();
            for (int k = 0; k < n; k++) {
                const uint32_t r = ordering[(size_t)l * n + k];
                bins[(size_t)(__builtin_popcount(r / M) & 1) * M + r % M].push_back(k);
            }
            for (int e = 0; e < EPT; e++)
                for (int sg = 0; sg < 2; sg++)
                    for (int h = 0; h < TEAM / LG; h++) {  // one pool per (half-)warp
                        E.clear();
                        for (int q = 0; q < LG; q++)
                            for (int k : bins[(size_t)sg * M + fast_bin(logM, TEAM, e, h * LG + q)]) E.push_back(PoolEdge{q, k % LG, k, 0});
                        ps.run_best(E, 8);
                        stat[(size_t)l * 2] += 8;
                        stat[(size_t)l * 2 + 1] += PoolScheduler::cost(E, 8);
                        std::fill(used.begin(), used.end(), 0);
                        for (const PoolEdge &pe : E) used[(size_t)pe.step * 32 + pe.bank] = 1;
                        uint16_t *dst = pt.inv2.data() + ((((size_t)l * EPT + e) * 2 + sg) * TEAM + h * LG) * 8;
                        for (int t = 0; t < 8; t++) {  // idle slots read a zero word in a bank no real term of the step uses
                            int fb = 0;
                            for (int b = 0; b < LG; b++) if (!used[(size_t)t * 32 + b]) { fb = b; break; }
                            const int zero_word = n + ((fb - n) & (LG - 1));
                            for (int q = 0; q < LG; q++) dst[q * 8 + t] = (uint16_t)(zero_word * ESZ);
                        }
                        for (const PoolEdge &pe : E) dst[pe.lane * 8 + pe.step] = (uint16_t)(pe.id * ESZ);
                    }
        }
    });
    for (int l = 0; l < L; l++) { pt.fold_steps += stat[(size_t)l * 2]; pt.fold_wavefronts += stat[(size_t)l * 2 + 1]; }
    const int G = L / 8;
    pt.fwd2.assign((size_t)G * n * 8, 0);
    std::vector<long> gstat((size_t)G * 2, 0);
    parallel_for(G, [&](int g0, int g1) {
        PoolScheduler ps;
        std::vector<PoolEdge> E;
        for (int g = g0; g < g1; g++)
            for (int k0 = 0; k0 < n; k0 += LG) {
                E.clear();
                const int nl = (n - k0 < LG) ? n - k0 : LG;
                for (int q = 0; q < nl; q++)
                    for (int i = 0; i < 8; i++) {
                        const uint32_t r = ordering[(size_t)(g * 8 + i) * n + k0 + q];
                        const uint32_t lo = r % M, sg = __builtin_popcount(r / M) & 1;
                        const uint32_t word = fq_word(logM, lo), off = (uint32_t)i * 4096u + word * (uint32_t)ESZ;
                        E.push_back(PoolEdge{q, (int)(word % (uint32_t)LG), (int)(off | (sg << 15)), 0});
                    }
                ps.run_best(E, 8);
                gstat[(size_t)g * 2] += 8;
                gstat[(size_t)g * 2 + 1] += PoolScheduler::cost(E, 8);
                for (const PoolEdge &pe : E) pt.fwd2[((size_t)g * n + k0 + pe.lane) * 8 + pe.step] = (uint16_t)pe.id;
            }
    });
    for (int g = 0; g < G; g++) { pt.gather_steps += gstat[(size_t)g * 2]; pt.gather_wavefronts += gstat[(size_t)g * 2 + 1]; }
}

int launch_amp2(const sb_operator *op, const AmpArgs &a, int B, int f64, cudaStream_t st);  // amp2.cu

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
    op->fwd = nullptr; op->inv = nullptr; op->fwd8 = nullptr; op->invq = nullptr; op->fwdq = nullptr;
    {   // FAST-mode tables (scheduled gathers)
        FastTables ft;
        build_fast_tables(ordering, L, M, n, op->Hp, ft);
        op->qok = ft.qok; op->qpre = ft.qpre; op->qneg = ft.qneg; op->PW = ft.PW; op->GQ = ft.GQ;
        if (ft.qok) {
            cudaError_t q1 = cudaMalloc(&op->invq, ft.invq.size() * 2), q2 = cudaSuccess;
            if (q1 == cudaSuccess) q1 = cudaMemcpy(op->invq, ft.invq.data(), ft.invq.size() * 2, cudaMemcpyHostToDevice);
            if (ft.GQ > 0) {
                q2 = cudaMalloc(&op->fwdq, ft.fwdq.size() * 2);
                if (q2 == cudaSuccess) q2 = cudaMemcpy(op->fwdq, ft.fwdq.data(), ft.fwdq.size() * 2, cudaMemcpyHostToDevice);
            }
            if (q1 != cudaSuccess || q2 != cudaSuccess) {
                cudaFree(op->invq); cudaFree(op->fwdq); free(hf); free(hi); free(h8); delete op;
                return fail(SB_ENOMEM, "sb_operator_create: cudaMalloc (fast tables) failed%s", "");
            }
        }
    }
    op->p2ok = 0; op->inv2 = nullptr; op->fwd2 = nullptr; op->inv2d = nullptr; op->fwd2d = nullptr; op->p2d_state = 0;
    op->h_ordering = nullptr;
    if (op->qok) {  // pair-kernel tables
        PairTables pt;
        build_pair_tables(ordering, L, M, n, op->H, pt);
        if (pt.ok) {
            cudaError_t q1 = cudaMalloc(&op->inv2, pt.inv2.size() * 2), q2 = cudaMalloc(&op->fwd2, pt.fwd2.size() * 2);
            if (q1 == cudaSuccess) q1 = cudaMemcpy(op->inv2, pt.inv2.data(), pt.inv2.size() * 2, cudaMemcpyHostToDevice);
            if (q2 == cudaSuccess) q2 = cudaMemcpy(op->fwd2, pt.fwd2.data(), pt.fwd2.size() * 2, cudaMemcpyHostToDevice);
            if (q1 != cudaSuccess || q2 != cudaSuccess) {
                cudaFree(op->inv2); cudaFree(op->fwd2); cudaFree(op->invq); cudaFree(op->fwdq); free(hf); free(hi); free(h8); delete op;
                return fail(SB_ENOMEM, "sb_operator_create: cudaMalloc (pair tables) failed%s", "");
            }
            op->p2ok = 1;
            // the fp64 tables of SB_AMP_F64 are built on first use (ensure_f64_tables): keep the ordering
            op->h_ordering = (uint32_t *)malloc((size_t)L * n * sizeof(uint32_t));
            if (op->h_ordering) memcpy(op->h_ordering, ordering, (size_t)L * n * sizeof(uint32_t));
        }
    }
    cudaError_t e1 = cudaMalloc(&op->fwd, nf * 2), e2 = cudaMalloc(&op->inv, ni * 2), e3 = cudaMalloc(&op->fwd8, n8 * 2);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(op->fwd); cudaFree(op->inv); cudaFree(op->fwd8); cudaFree(op->invq); cudaFree(op->fwdq); cudaFree(op->inv2); cudaFree(op->fwd2);
        free(hf); free(hi); free(h8); delete op;
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
    cudaFree(op->invq);
    cudaFree(op->fwdq);
    cudaFree(op->inv2);
    cudaFree(op->fwd2);
    cudaFree(op->inv2d);
    cudaFree(op->fwd2d);
    free(op->h_ordering);
    delete op;
}

// Test hook (no GPU needed) for the pair-kernel tables: every (bin, sign half) of every section lists exactly its
// rows, every (8-section group, row) exactly its 8 (slot, column, sign) terms.  stats[0..3] as sb_fast_tables_check.
// Returns 0 = verified, 1 = the shape has no pair tables, < 0 = error.
static int pair_tables_check(const uint32_t *ordering, int L, int M, int n, int f64, long *stats) {
    if (!ordering || L <= 0 || n <= 0 || M < 2 || (M & (M - 1)) || M > 1024 || n >= 65534)
        return fail(SB_EINVAL, "sb_pair_tables_check: bad shape%s (M=%ld)", "", M);
    int w = 1;
    while (w < (M + 1 > n + 1 ? M + 1 : n + 1)) w <<= 1;
    PairTables pt;
    build_pair_tables(ordering, L, M, n, w / M, pt, f64);
    if (stats) { stats[0] = pt.fold_steps; stats[1] = pt.fold_wavefronts; stats[2] = pt.gather_steps; stats[3] = pt.gather_wavefronts; }
    if (!pt.ok) return 1;
    const int esh = f64 ? 3 : 2, emask = (1 << esh) - 1;
    std::vector<int> seen(n);
    for (int l = 0; l < L; l++) {
        std::fill(seen.begin(), seen.end(), 0);
        for (int e = 0; e < 16; e++)
            for (int sg = 0; sg < 2; sg++)
                for (int q = 0; q < 32; q++)
                    for (int t = 0; t < 8; t++) {
                        const int o = pt.inv2[((((size_t)l * 16 + e) * 2 + sg) * 32 + q) * 8 + t];
                        if (o & emask) return fail(SB_EINVAL, "pair tables: unaligned fold offset%s (%ld)", "", o);
                        const int k = o >> esh;
                        if (k >= n && k < n + 32) continue;  // zero word
                        if (k >= n + 32) return fail(SB_EINVAL, "pair tables: fold offset out of range%s (%ld)", "", o);
                        const uint32_t r = ordering[(size_t)l * n + k];
                        if ((int)(r % M) != fast_bin(9, 32, e, q) || (__builtin_popcount(r / M) & 1) != sg || seen[k]++)
                            return fail(SB_EINVAL, "pair tables: wrong fold term%s (section %ld)", "", l);
                    }
        for (int k = 0; k < n; k++) if (seen[k] != 1) return fail(SB_EINVAL, "pair tables: missing fold term%s (section %ld)", "", l);
    }
    for (int g = 0; g < L / 8; g++)
        for (int k = 0; k < n; k++) {
            unsigned mask = 0;
            for (int t = 0; t < 8; t++) {
                const uint32_t e = pt.fwd2[((size_t)g * n + k) * 8 + t];
                const uint32_t sg = e >> 15, off = e & 0x7FFFu, slot = off >> 12, word = (off & 4095u) >> esh;
                if ((off & (uint32_t)emask) || word >= 512) return fail(SB_EINVAL, "pair tables: bad gather offset%s (%ld)", "", e);
                const uint32_t lo = ((word >> 1) & 15u) * 32u + 2u * (word >> 5) + (word & 1u);  // inverse of fq_word
                const uint32_t r = ordering[(size_t)(g * 8 + slot) * n + k];
                if (r % M != lo || (uint32_t)(__builtin_popcount(r / M) & 1) != sg)
                    return fail(SB_EINVAL, "pair tables: wrong gather term%s (group %ld)", "", g);
                mask |= 1u << slot;
            }
            if (mask != 0xFFu) return fail(SB_EINVAL, "pair tables: missing gather term%s (group %ld)", "", g);
        }
    return SB_OK;
}

extern "C" int sb_pair_tables_check(const uint32_t *ordering, int L, int M, int n, long *stats) {
    return pair_tables_check(ordering, L, M, n, 0, stats);
}
// the same check for the fp64 tables of SB_AMP_F64 (pools of 16 lanes x 16 eight-byte banks)
extern "C" int sb_pair_tables_check_f64(const uint32_t *ordering, int L, int M, int n, long *stats) {
    return pair_tables_check(ordering, L, M, n, 1, stats);
}

// Test hook (no GPU needed): builds the FAST-mode tables on the host and verifies that they are a reordering of
// the operator -- every bin of every section lists exactly its (k, sign) terms, every (16-section chunk, row)
// lists exactly its 16 (slot, lo, sign) terms.  stats[0..3] = fold steps, fold wavefronts, gather steps, gather
// wavefronts of the bank-conflict model (wavefronts / steps = 1 means conflict-free).
extern "C" int sb_fast_tables_check(const uint32_t *ordering, int L, int M, int n, long *stats) {
    if (!ordering || L <= 0 || n <= 0 || M < 2 || (M & (M - 1)) || M > 1024 || n >= 65534)
        return fail(SB_EINVAL, "sb_fast_tables_check: bad shape%s (M=%ld)", "", M);
    int w = 1;
    while (w < (M + 1 > n + 1 ? M + 1 : n + 1)) w <<= 1;
    const int H = w / M, Hp = H < 16 ? 16 : H, logM = ilog2(M), SBQ = sign_bit(logM);
    FastTables ft;
    build_fast_tables(ordering, L, M, n, Hp, ft);
    if (stats) { stats[0] = ft.fold_steps; stats[1] = ft.fold_wavefronts; stats[2] = ft.gather_steps; stats[3] = ft.gather_wavefronts; }
    if (!ft.qok) return 1;
    const int osh = ft.qpre ? 2 : 0, TEAM = team_lanes(M), EPT = M / TEAM, NH = Hp / 8, PW = ft.PW;
    std::vector<int> seen(n);
    for (int l = 0; l < L; l++) {
        std::fill(seen.begin(), seen.end(), 0);
        for (int eq = 0; eq < M; eq++)
            for (int t = 0; t < Hp; t++) {
                const int j = fast_bin(logM, TEAM, eq / TEAM, eq % TEAM);
                const int o = ft.invq[((((size_t)l * EPT + eq / TEAM) * NH + (t >> 3)) * TEAM + eq % TEAM) * 8 + (t & 7)] >> osh;
                if (o >= n && o < n + 32) continue;  // zero word
                const int neg = o >= ft.qneg, k = neg ? o - ft.qneg : o;
                if (k < 0 || k >= n) return fail(SB_EINVAL, "fast tables: fold offset out of range%s (%ld)", "", o);
                const uint32_t r = ordering[(size_t)l * n + k];
                if ((int)(r % M) != j || (__builtin_popcount(r / M) & 1) != neg || seen[k]++)
                    return fail(SB_EINVAL, "fast tables: wrong fold term%s (section %ld)", "", l);
            }
        for (int k = 0; k < n; k++) if (seen[k] != 1) return fail(SB_EINVAL, "fast tables: missing fold term%s (section %ld)", "", l);
    }
    for (int g = 0; g < ft.GQ; g++)
        for (int k = 0; k < n; k++) {
            unsigned mask = 0;
            for (int t = 0; t < PW; t++) {
                const uint32_t off = ft.fwdq[(((size_t)g * (PW / 8) + (t >> 3)) * n + k) * 8 + (t & 7)];
                const uint32_t region = off / (2u << SBQ), in = off % (2u << SBQ);
                const uint32_t sg = in >> SBQ, rest = in & ((1u << SBQ) - 1), slot = region * 8 + rest / ((uint32_t)M << 2);
                uint32_t lo = (rest % ((uint32_t)M << 2)) >> 2;
                if (logM == 9) lo = ((lo >> 1) & 15u) * 32u + 2u * (lo >> 5) + (lo & 1u);  // inverse of fq_word
                if (slot >= (uint32_t)PW || (off & 3)) return fail(SB_EINVAL, "fast tables: bad gather offset%s (%ld)", "", off);
                const uint32_t r = ordering[(size_t)(g * PW + slot) * n + k];
                if (r % M != lo || (uint32_t)(__builtin_popcount(r / M) & 1) != sg)
                    return fail(SB_EINVAL, "fast tables: wrong gather term%s (chunk %ld)", "", g);
                mask |= 1u << slot;
            }
            if (mask != (PW == 16 ? 0xFFFFu : 0xFFu)) return fail(SB_EINVAL, "fast tables: missing gather term%s (chunk %ld)", "", g);
        }
    return SB_OK;
}

static AmpArgs base_args(const sb_operator *op, const int *sections, const int *nsec) {
    AmpArgs a;
    memset(&a, 0, sizeof(a));
    a.fwd = op->fwd; a.fwd8 = op->fwd8; a.inv = op->inv; a.sections = sections; a.nsec = nsec;
    a.invq = op->invq; a.fwdq = op->fwdq; a.qneg = op->qneg; a.PW = op->PW;
    a.L = op->L; a.n = op->n; a.Hp = op->Hp; a.NB = op->NB;
    return a;
}

#ifdef SB_PHASE_CLOCKS
static unsigned long long *g_dbg = nullptr;
// experiment builds only: read (and clear) the per-phase cycle counters of the AMP kernel
extern "C" int sb_phase_cycles_read(unsigned long long *out16) {
    if (!g_dbg) { memset(out16, 0, 16 * sizeof(unsigned long long)); return SB_OK; }
    SB_CUDA(cudaMemcpy(out16, g_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    SB_CUDA(cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long)));
    return SB_OK;
}
#endif

// Diagnostic switch (A/B timing and the pair-vs-single parity test): 0 makes FAST mode use the one-codeword-per-CTA
// kernel for every shape.  Returns the previous setting.  Initial value: 1, or 0 when SB_AMP_PAIR=0 is set.
static std::atomic<int> g_pair_on{[] { const char *e = getenv("SB_AMP_PAIR"); return (e && e[0] == '0') ? 0 : 1; }()};
extern "C" int sb_amp_pair_enable(int on) { return g_pair_on.exchange(on ? 1 : 0); }

// fp64 tables of the warp-specialised kernel, built once per operator on the first SB_AMP_F64 call
static std::mutex g_f64_mu;
static int ensure_f64_tables(sb_operator *op) {
    std::lock_guard<std::mutex> lk(g_f64_mu);
    if (op->p2d_state) return op->p2d_state > 0 ? SB_OK : 1;
    if (!op->p2ok || !op->h_ordering) { op->p2d_state = -1; return 1; }
    PairTables pt;
    build_pair_tables(op->h_ordering, op->L, op->M, op->n, op->H, pt, 1);
    if (!pt.ok) { op->p2d_state = -1; return 1; }
    cudaError_t e1 = cudaMalloc(&op->inv2d, pt.inv2.size() * 2), e2 = cudaMalloc(&op->fwd2d, pt.fwd2.size() * 2);
    if (e1 == cudaSuccess) e1 = cudaMemcpy(op->inv2d, pt.inv2.data(), pt.inv2.size() * 2, cudaMemcpyHostToDevice);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(op->fwd2d, pt.fwd2.data(), pt.fwd2.size() * 2, cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {  // no tables: this and later F64 calls run the STRICT kernel
        cudaFree(op->inv2d); cudaFree(op->fwd2d);
        op->inv2d = nullptr; op->fwd2d = nullptr;
        op->p2d_state = -1;
        (void)cudaGetLastError();
        return 1;
    }
    free(op->h_ordering);
    op->h_ordering = nullptr;
    op->p2d_state = 1;
    return SB_OK;
}

extern "C" int sb_amp_batch(const sb_operator *op, const double *y, const double *Pl, const double *beta0,
                            const int *sections, const int *nsec, int B, int T, int mode, double *beta, int *iters,
                            int *n_exec, unsigned *flags, double *tau2_trace, double *scratch, void *stream) {
    if (!op || !y || !Pl || !beta || !iters || !n_exec || !flags || B < 0 || T < 0)
        return fail(SB_EINVAL, "sb_amp_batch: null argument%s", "");
    if (mode != SB_AMP_STRICT && mode != SB_AMP_FAST && mode != SB_AMP_F64) return fail(SB_EINVAL, "sb_amp_batch: unknown mode%s %ld", "", mode);
    if ((sections == nullptr) != (nsec == nullptr)) return fail(SB_EINVAL, "sb_amp_batch: sections and nsec go together%s", "");
    if ((mode == SB_AMP_FAST && op->qok && !scratch) || (mode == SB_AMP_F64 && op->p2ok && !scratch))
        return fail(SB_EINVAL, "sb_amp_batch: FAST / F64 mode needs a [2][B][n] scratch%s", "");
    if (B == 0) return SB_OK;
    AmpArgs a = base_args(op, sections, nsec);
    a.zscratch = scratch;
    a.y = y; a.Pl = Pl; a.beta0 = beta0; a.beta = beta; a.tau2_trace = tau2_trace;
    a.iters = iters; a.n_exec = n_exec; a.flags = flags; a.T = T;
#ifdef SB_PHASE_CLOCKS
    if (!g_dbg) { SB_CUDA(cudaMalloc(&g_dbg, 16 * sizeof(unsigned long long))); SB_CUDA(cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long))); }
    a.dbg = g_dbg;
#endif
    // FAST, all sections active, M = 512: the warp-specialised two-codeword kernel (amp2.cu)
    if (mode == SB_AMP_FAST && op->p2ok && sections == nullptr && g_pair_on.load()) return launch_amp2(op, a, B, 0, (cudaStream_t)stream);
    // F64: the same kernel structure on fp64 values, one codeword per CTA; shapes / section lists it does not cover
    // run the order-preserving STRICT kernel (same arithmetic type and stop rule)
    if (mode == SB_AMP_F64 && op->p2ok && sections == nullptr && g_pair_on.load()) {
        const int rc = ensure_f64_tables(const_cast<sb_operator *>(op));
        if (rc == SB_OK) return launch_amp2(op, a, B, 1, (cudaStream_t)stream);
        if (rc < 0 || rc > 1) return rc;
    }
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
