/*
 * sparc_b200.h -- C ABI of libsparc_b200.so, the B200 (sm_100a) drop-in for the
 * SPARC-AMP + outer-LDPC decode path of Spimp/sparc_ldpc.
 *
 * Two groups of entry points:
 *
 *  (1) The reference's own FFI, unchanged (ldpc/src/c_ldpc.c, bound by
 *      ldpc/py/ldpc.py:859-943 through ctypes.CDLL('./bin/c_ldpc.so')):
 *      sumprod, sumprod2, minsum, Lxor, Lxfb -- HOST pointers, one codeword.
 *      Pointing ldpc.py at this library instead of c_ldpc.so needs no other change
 *      (see INTEGRATION.md).
 *
 *  (2) Batched, device-pointer entry points (prefix sb_) that replace the Python /
 *      numpy hot path (ldpc/sparc_ldpc.py:32-356, ldpc/amp_exit.py:56-122,:272-305)
 *      and the per-codeword ctypes call.  Every array argument of an sb_*_batch
 *      function is a DEVICE pointer unless its name ends in _host; `stream` is a
 *      cudaStream_t passed as void* (NULL = legacy default stream).  Layouts are
 *      row-major, codeword-major: x[b][i].
 *
 * Conventions: all functions return >= 0 on success and a negative SB_E* code on
 * failure, never throw, never call exit().  sb_last_error() returns a thread-local
 * message for the last failure.  `long` is 64-bit (LP64), as the reference assumes
 * (numpy int64 passed as c_long, ldpc.py:870-872).
 */
#ifndef SPARC_B200_H
#define SPARC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_EINVAL (-2)   /* bad argument / unsupported shape            */
#define SB_ECUDA (-3)    /* CUDA runtime error (see sb_last_error())     */
#define SB_ENOMEM (-1)   /* allocation failure (the reference's only error code, c_ldpc.c:165-166) */

#define SB_MAX_ITCOUNT 200 /* ldpc/src/c_ldpc.c:7 */

/* decode rules of sb_bp_batch */
#define SB_BP_SUMPROD2 0 /* c_ldpc.c:138-206 (default of ldpc.py:855)  */
#define SB_BP_SUMPROD 1  /* c_ldpc.c:32-113  (tanh / atanh rule)        */
#define SB_BP_MINSUM 2   /* c_ldpc.c:339-381 (without the :364 indexing bug) */
#define SB_BP_SUMPROD2_FAST 3 /* sumprod2 with the two log(1+exp(-|x|)) correction terms of every Lxor
                                 (c_ldpc.c:246-247) evaluated in single precision; sign-min term, variable-node
                                 sums, app and the stop test in fp64.  app agrees with SB_BP_SUMPROD2 to ~1e-6,
                                 iteration counts and hard decisions are equal except for near-ties */

/* per-codeword status bits written by sb_amp_batch into flags[b] */
#define SB_AMP_STOPPED 1u   /* tau == last_tau fired (sparc_ldpc.py:204)                        */
#define SB_AMP_REF_NAN 2u   /* the reference's global-max softmax (sparc_ldpc.py:216-219) left fp64's normal range
                               on this codeword: some section's maximum lay > 708.4 below the global one (sums of
                               subnormals, relative error ~1e-5) or > 745.1 below it (0/0 = NaN); the kernel itself
                               computes every section accurately */

const char *sb_last_error(void);
int sb_version(void);
/* number of kernels launched by this library since load / since the last reset (all threads) */
long sb_launch_count(void);
void sb_launch_count_reset(void);

/* ------------------------------------------------------------------ (1) reference FFI
 * Replaces ldpc/src/c_ldpc.c:32 / :138 / :339 / :234 / :294 symbol for symbol.
 * Return value of the decoders: iterations used (0..200), -1 on allocation failure.
 * `app` (length Nv) is fully overwritten; the caller owns every buffer. */
int sumprod(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app);
int sumprod2(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app);
int minsum(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg, double *app,
           double correction_factor);
double Lxor(double L1, double L2, int corr_flag);
double Lxfb(double *L, long dc, int corr_flag); /* overwrites L[0..dc) with extrinsics, returns the total */

/* ------------------------------------------------------------------ (2a) Tanner graph handle
 * Built once per code from the arrays ldpc.py:694-786 (prepare_decoder) returns. */
typedef struct sb_graph sb_graph;
int sb_graph_create(const long *vdeg_host, const long *cdeg_host, const long *intrlv_host, int Nv, int Nc, int Nmsg,
                    sb_graph **out);
void sb_graph_destroy(sb_graph *g);

/* Batched flooding BP, one CTA per codeword, messages resident in shared memory.
 * ch[B][Nv] in, app[B][Nv] out, it[B] out (iteration index at which every check was
 * satisfied, or max_it).  Replaces ldpc.py:855-930 -> c_ldpc.c:138-206. */
int sb_bp_batch(const sb_graph *g, int rule, const double *ch, int B, double *app, int *it, int max_it,
                double minsum_factor, void *stream);

/* Measured instruction ceiling of the check-node function for `rule` on the current device: Lxor evaluations per
 * second of a register-only loop on all SMs (no memory traffic).  bench.py reports the BP kernel's achieved Lxor/s
 * against it (c_ldpc.c:234-251 evaluated 3 (dc - 2) times per check and iteration, :294-314). */
int sb_bp_lxor_peak(int rule, double *lxor_per_s);

/* ------------------------------------------------------------------ (2b) SPARC design operator
 * Block sub-sampled Walsh-Hadamard operator of sparc_ldpc.py:81-147, built from the
 * (L, n) uint32 `ordering` table sparc_transforms returns (host pointer). */
typedef struct sb_operator sb_operator;
int sb_operator_create(const uint32_t *ordering_host, int L, int M, int n, sb_operator **out);
void sb_operator_destroy(sb_operator *op);
/* Host-only self check of the FAST-mode lookup tables (no GPU needed): rebuilds them from `ordering_host` and
 * verifies that they are a reordering of the operator of sparc_ldpc.py:110-134 (every bin lists exactly its
 * (row, sign) terms, every (16-section chunk, row) exactly its 16 (section, column, sign) terms).
 * stats[4] (may be NULL) = fold steps, fold shared-memory wavefronts, gather steps, gather wavefronts of the
 * bank-conflict model.  Returns 0 = verified, 1 = shape has no FAST tables, <0 = error. */
int sb_fast_tables_check(const uint32_t *ordering_host, int L, int M, int n, long *stats);
/* The same self check for the tables of the two-codewords-per-CTA FAST kernel (M = 512, w/M <= 16, n <= 4608,
 * L % 8 == 0; csrc/amp2.cu): every (bin, sign half) lists exactly its rows, every (8-section group, row) its 8
 * (section, column, sign) terms.  Returns 0 = verified, 1 = the shape has no pair tables, <0 = error. */
int sb_pair_tables_check(const uint32_t *ordering_host, int L, int M, int n, long *stats);
int sb_pair_tables_check_f64(const uint32_t *ordering_host, int L, int M, int n, long *stats); /* the fp64 tables of SB_AMP_F64 */

/* pyfht.fht_inplace (sparc_ldpc.py:14-29, :69, :76): HOST pointer, N a power of two, transformed in place. */
int sb_fht_inplace_host(double *x, long N);

/* out[b][0..n) = A_S beta[b] where S = the codeword's section list (sparc_ldpc.py:143-144;
 * sparc_transforms_shorter :154-168 when `sections` is given).  beta[b] is compact:
 * nsec[b]*M entries.  sections == NULL: all L sections in order; nsec == NULL: L. */
int sb_Ab_batch(const sb_operator *op, const double *beta, const int *sections, const int *nsec, int B, double *out,
                void *stream);
/* out[b][0..nsec*M) = A_S^T z[b]  (sparc_ldpc.py:145-146) */
int sb_Az_batch(const sb_operator *op, const double *z, const int *sections, const int *nsec, int B, double *out,
                void *stream);

/* out[b][k] = (y ? y[b][k] : 0) + sign * sum_l sqrt(n*Pl[l]) * A[k, l*M + idx[b][l]]   over sections with idx >= 0.
 * The SPARC encoder (sparc_ldpc.py:436-439, sign=+1, y=NULL) and the peeling subtraction
 * (sparc_ldpc.py:508-518, amp_exit.py:111-112, sign=-1). */
int sb_onehot_apply_batch(const sb_operator *op, const int *idx, const double *Pl, const double *y, double sign, int B,
                          double *out, void *stream);

/* AMP decoder, sparc_ldpc.py:189-222 / amp_test.py:14-50, one persistent CTA per codeword.
 *   y[B][n]; Pl[L] (indexed by section id); beta0[B][L*M] or NULL (zero start);
 *   sections[B][L] / nsec[B] or NULL (all sections): active-section lists, beta is compact;
 *   beta[B][L*M] out; iters[B] out = the `t` amp_test returns; n_exec[B] out = iterations
 *   actually executed; flags[B] out (SB_AMP_*); tau2_trace[B][T] or NULL. */
#define SB_AMP_STRICT 0 /* fp64 everywhere; the two operator products inside the loop add in the reference's order (the
                           kernels behind sb_Ab_batch / sb_Az_batch, which ARE bit-identical to the reference).  The decode as a
                           whole is not bit-faithful -- section maximum instead of the global one, multiplication by 1/sqrt(n),
                           CUDA's exp, FMA contraction -- and agrees with the reference to ~1e-12 per iteration              */
#define SB_AMP_FAST 1   /* z and FHT(beta) are gathered from 32-bit fixed-point copies (27 bits below their
                           power-of-two ceiling), adds of the two gathers are exact integer adds; FHT, softmax,
                           z update and tau^2 stay fp64.  beta / tau^2 agree with STRICT to ~1e-8 relative.  The
                           stop rule becomes |tau - last_tau| <= 2^-27 tau (tau cannot reach an exact fixed point
                           above the quantisation floor), so `iters` is smaller than in STRICT mode.                */
#define SB_AMP_F64 2    /* fp64 everywhere and the reference's exact-equality stop rule, like STRICT, but the two gathers add
                           their terms in an order chosen for conflict-free shared-memory access (as FAST does with its
                           integers): beta / tau^2 differ from STRICT by fp64 summation-order noise (~1e-15 relative per
                           iteration).  M = 512, w/M <= 16, n <= 4608, L % 8 == 0, all sections active: the warp-specialised
                           kernel of csrc/amp2.cu on fp64 values, one codeword per CTA; any other call runs STRICT.       */
/* FAST mode has two kernels: one CTA per codeword (any shape) and, for M = 512 with w/M <= 16, n <= 4608, L % 8 == 0
 * and all sections active, a warp-specialised kernel that decodes two codewords per CTA (csrc/amp2.cu; same
 * arithmetic, A beta / A^T z bit-identical).  sb_amp_pair_enable(0) forces the first for every shape (A/B timing,
 * parity tests); returns the previous setting. */
int sb_amp_pair_enable(int on);
int sb_amp_batch(const sb_operator *op, const double *y, const double *Pl, const double *beta0, const int *sections,
                 const int *nsec, int B, int T, int mode, double *beta, int *iters, int *n_exec, unsigned *flags,
                 double *tau2_trace, double *scratch /* [2][B][n] work space, required in FAST mode: z and, for n > 8191,
                                                        the A beta accumulator */, void *stream);

/* ------------------------------------------------------------------ (2c) section <-> bit handoff
 * p = sp2bp(beta / sqrt(n*Pl)) (sparc_ldpc.py:257-281, :657); llr = nan_to_num(log(1-p) - log(p)) (:667-669).
 * beta rows have stride `beta_stride` doubles.  Codeword b processes a list of sections:
 *   sections == NULL: entries i = 0..count-1, section id s = first_sec + i (indexes Pl),
 *                     beta block (beta_first + i), output bits at (out_first + i)*logM ..;
 *   sections != NULL: entries i = 0..nsec[b]-1, s = sections[b][i] (row stride L_stride ints),
 *                     beta block (beta_first + i), output bits at s*logM .. (sparc_ldpc.py:1031-1034).
 * p and llr rows have stride `out_stride` doubles; p may be NULL. */
int sb_sp2bp_llr_batch(const double *beta, long beta_stride, int beta_first, const int *sections, const int *nsec,
                       int L_stride, int first_sec, int out_first, int count, int M, int n, const double *Pl, int B,
                       double *p, double *llr, long out_stride, void *stream);

/* Section priors / posteriors from bit information (sparc_ldpc.py:283-314, :685-696).  in[B][ls*logM] holds the
 * LDPC a-posteriori LLRs (bit posterior 1/(1+exp(app)), :685) or, with SB_PRIOR_FROM_PROB, bit probabilities.
 * Protected sections [L-ls, L) get bp2sp(bits); unprotected sections get beta_prev/sqrt(n*Pl) (:657).  With
 * SB_PRIOR_SCALE every section is then multiplied by sqrt(n*Pl) to give the next AMP initialisation (:696). */
#define SB_PRIOR_SCALE 1
#define SB_PRIOR_FROM_PROB 2
int sb_bp2sp_prior_batch(const double *in, int ls, const double *beta_prev, int L, int M, int n, const double *Pl,
                         int mode, int B, double *out, void *stream);

/* Denoiser of amp() alone (sparc_ldpc.py:214-219), for design matrices that are not the structured operator
 * (dense Gaussian A: the two products are plain GEMMs done by the caller).  s[B][L*M] = beta + A^T z,
 * tau2[B]; beta[B][L*M] out, sumsq[B][L] out = sum(beta_l^2) per section (0 for inactive codewords);
 * active[B] (bytes) or NULL selects codewords. */
int sb_section_softmax_batch(const double *s, const double *Pl, const double *tau2, const unsigned char *active, int L,
                             int M, int n, int B, double *beta, double *sumsq, void *stream);

/* ------------------------------------------------------------------ (2c) dense (Gaussian) design matrix
 * amp() of sparc_ldpc.py:189-222 with Ab = A @ beta, Az = A.T @ z for an explicit matrix A [n][L*M] (the
 * reference takes the operator as two closures, :189; BASELINE configs[0]).  Both products run as hand-written
 * tcgen05 / TMA GEMMs over the batch with a bf16x3 split of every operand (FP32 emulation: operands carried to
 * 2^-27, six bf16 MMA passes, fp32 TMEM accumulation over at most 1024 values of k, fp64 across chunks); the
 * AMP state (beta, z, tau^2, softmax, Onsager term) stays fp64.  A_dev: DEVICE pointer, row-major fp64. */
typedef struct sb_dense sb_dense;
int sb_dense_create(const double *A_dev, int n, int LM, sb_dense **out);
void sb_dense_destroy(sb_dense *d);
/* out[b] = A x[b]  (transpose = 0: x [B][LM] -> out [B][n])  or  A^T x[b]  (transpose = 1: x [B][n] -> out [B][LM]) */
int sb_dense_apply_batch(sb_dense *d, int transpose, const double *x, int B, double *out, void *stream);
/* Batched AMP decode; device pointers: y [B][n], Pl [L], beta0 [B][L*M] or NULL, beta [B][L*M] out,
 * iters / n_exec / flags [B] out (as sb_amp_batch), tau2_trace [B][T] or NULL (NaN where not executed).
 * Stop rule: tau == last_tau (sparc_ldpc.py:204) or |tau - last_tau| <= 2^-27 tau. */
int sb_dense_amp_batch(sb_dense *d, const double *y, const double *Pl, const double *beta0, int L, int M, int B, int T,
                       double *beta, int *iters, int *n_exec, unsigned *flags, double *tau2_trace, void *stream);

/* Column-sharded dense matrix (for an A that exceeds one GPU's HBM): the handle holds the columns of L_local
 * sections; Pl_local / beta0_local / beta_local refer to them, P_total = sum of Pl over ALL sections; y, z and
 * tau^2 are replicated on every rank.  A^T z, the softmax and |beta|^2 are local; once per iteration the library
 * calls `allreduce(ctx, xbuf, B*n + B, stream)`, which must sum xbuf (partial A beta [B][n] followed by |beta|^2
 * [B], device memory owned by the caller) over all ranks in place, ordered on `stream`, and return 0.  With
 * NCCL: ncclAllReduce(xbuf, xbuf, count, ncclDouble, ncclSum, comm, stream). */
typedef int (*sb_allreduce_fn)(void *ctx, double *buf_dev, long count, void *stream);
int sb_dense_amp_batch_sharded(sb_dense *d, const double *y, const double *Pl_local, double P_total,
                               const double *beta0_local, int L_local, int M, int B, int T, double *beta_local, int *iters,
                               int *n_exec, unsigned *flags, double *tau2_trace, double *xbuf, sb_allreduce_fn allreduce,
                               void *ctx, void *stream);

/* The same column-sharded decode with the exchange done over NVLink PEER MEMORY instead of a collective library:
 * every rank owns a receive area of 2 * world * (B*n + B) doubles and `world` 64-bit flag words (zero-initialised),
 * both mapped by all peers (CUDA IPC across processes, plain pointers inside one process).  Per iteration one kernel
 * folds the K slices of this rank's partial A beta and pushes it, with |beta|^2, into slot `rank` of every peer's
 * area; the last CTA publishes the epoch in every peer's flag word (system-scope release); the residual kernel adds
 * the slots in rank order after all local flags have arrived, so z stays bit-identical on every rank.
 * slots[r] / flags[r]: rank r's area / flag words AS MAPPED IN THIS PROCESS (r = rank: the local ones).
 * *epoch_io: exchange counter, 0 before the first call, carried from call to call (all ranks make the same calls);
 *            written back on EVERY return path.  After a failed call (SB_ECUDA: a peer did not arrive) the exchange
 *            is out of step with its peers: free and re-allocate the areas collectively before using it again.
 * timeout_ms: a peer that does not arrive in time makes the call fail with SB_ECUDA instead of hanging. */
#define SB_P2P_MAX 8
typedef struct {
    int rank, world, timeout_ms;
    double *slots[SB_P2P_MAX];
    unsigned long long *flags[SB_P2P_MAX];
    long slot_doubles; /* doubles per slot the areas were ALLOCATED for (area = 2 * world * slot_doubles): the slot and
                          parity strides, fixed for the life of the exchange, so that calls with different B never
                          overlap an area a slower peer may still be reading; every call needs B*n + B <= slot_doubles */
} sb_p2p;
int sb_enable_peer_access(int peer_device); /* cudaDeviceEnablePeerAccess from the current device */
/* receive areas shared between processes: cudaMalloc + cudaIpcGetMemHandle on the owner (zero-initialised),
 * cudaIpcOpenMemHandle(cudaIpcMemLazyEnablePeerAccess) on every peer WITH ITS OWN DEVICE CURRENT */
int sb_p2p_alloc(long bytes, void **dev_ptr, unsigned char *handle64);
int sb_p2p_open(const unsigned char *handle64, void **dev_ptr);
int sb_p2p_close(void *dev_ptr);
int sb_p2p_free(void *dev_ptr);
int sb_dense_amp_batch_p2p(sb_dense *d, const double *y, const double *Pl_local, double P_total,
                           const double *beta0_local, int L_local, int M, int B, int T, double *beta_local, int *iters,
                           int *n_exec, unsigned *flags, double *tau2_trace, double *xbuf, const sb_p2p *peers,
                           unsigned long long *epoch_io, void *stream);

/* idx[b][i] = argmax of section i of beta[b] (first maximum wins, sparc_ldpc.py:640-643) */
int sb_argmax_batch(const double *beta, long beta_stride, int count, int M, int B, int *idx, long idx_stride,
                    void *stream);
/* idx[b][i] = MSB-first index of hard decisions (llr < 0) (sparc_ldpc.py:317-341, :352, :672-674) */
int sb_llr2idx_batch(const double *llr, long llr_stride, int count, int M, int B, int *idx, long idx_stride,
                     void *stream);
/* errs[b] = sum_i popcount(a[b][i] ^ t[b][i]) (sparc_ldpc.py:650) */
int sb_count_errors_batch(const int *a, const int *t, int count, int B, int *errs, void *stream);

/* beta[b][l*M + idx[b][l]] = sqrt(n*Pl[l]), zero elsewhere (sparc_ldpc.py:840-843); idx < 0: all-zero section */
int sb_onehot_beta_batch(const int *idx, const double *Pl, int n, int L, int M, int B, double *beta, void *stream);

/* On-device input generation.  Systematic QC-LDPC encoder of ldpc.py:790-850: proto[Mp][Np]
 * (device, shift or -1), toff = the single odd-multiplicity shift (mod z) of protograph column Np-Mp (ldpc.py:823-835);
 * info[B][(Np-Mp)*z] bytes 0/1 -> x[B][Np*z] bytes 0/1. */
int sb_ldpc_encode_batch(const int *proto, int Mp, int Np, int z, int toff, const unsigned char *info, int B,
                         unsigned char *x, void *stream);
/* idx[b][i] = MSB-first value of bits[b][i*logM .. i*logM+logM) (bytes 0/1)   (bits2indices, sparc_ldpc.py:317-341) */
int sb_bits2idx_batch(const unsigned char *bits, long bits_stride, int count, int M, int B, int *idx, long idx_stride,
                      void *stream);

/* Threshold peel (amp_exit.py:85-116): post[B][L*M] section posteriors; sections l >= L-ls with exactly one
 * entry > threshold are hard decided.  hard_idx[B][L] = decided index or -1; act[B][L] = ascending list of the
 * remaining sections, nact[B] its length. */
int sb_threshold_peel_batch(const double *post, int L, int M, int ls, double threshold, int B, int *hard_idx,
                            int *act, int *nact, void *stream);

/* EXIT histograms (amp_exit.py:299-305): counts[b][0][.] for X == +1 and counts[b][1][.] for X == -1 of E[b][.]
 * over the nbins = nedges-1 bins delimited by edges[] (numpy.histogram semantics: left-closed, last bin closed). */
int sb_exit_hist_batch(const double *E, const int *X, int len, const double *edges, int nedges, int B,
                       long long *counts, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPARC_B200_H */
