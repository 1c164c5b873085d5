/*
 * oracle.c -- CPU restatement of the reference's native arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sparc_ldpc_b200/ may link, load or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, and only as the checker / the timed
 * CPU baseline.  Every function cites the reference lines it restates
 * (paths relative to /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks these functions
 * against golden vectors produced by the unmodified reference
 * (tests/golden/gen_golden.py) and, where oracle/_ref/c_ldpc.so has been built
 * from the reference's own ldpc/src/c_ldpc.c, against that library directly.
 *
 * The operators deliberately follow the reference's *literal* algorithm
 * (zero-pad to w, full w-point transform, gather) and not the M-point
 * Kronecker shortcut the CUDA path uses, so that the oracle is an independent
 * check of that shortcut.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- ldpc/sparc_ldpc.py:19-29  fht_inplace (the in-repo fallback that pins
 * pyfht's arithmetic): strides N/2 ... 1, (a, b) -> (a + b, a - b). ---------- */
void orc_fht_inplace(double *x, long N)
{
    for (long h = N >> 1; h; h >>= 1)
        for (long k = 0; k < N; k += 2 * h)
            for (long j = k; j < k + h; j++) {
                double a = x[j], b = x[j + h];
                x[j] = a + b;
                x[j + h] = a - b;
            }
}

/* ---- ldpc/sparc_ldpc.py:65-70 + :120-126  Ax of block_sub_fht:
 * per section zero-pad beta_l into the last m slots of a w-vector, transform,
 * gather ordering[l]; sections are accumulated in order l = 0..L-1.
 * `out` is NOT divided by sqrt(n) (sparc_transforms does that, :144). ------- */
int orc_block_Ax(const uint32_t *ordering, long L, long n, long m, long w,
                 const double *x, double *out)
{
    double *buf = (double *)malloc(sizeof(double) * (size_t)w);
    if (!buf) return -1;
    for (long k = 0; k < n; k++) out[k] = 0.0;
    for (long l = 0; l < L; l++) {
        memset(buf, 0, sizeof(double) * (size_t)w);
        memcpy(buf + (w - m), x + l * m, sizeof(double) * (size_t)m);
        orc_fht_inplace(buf, w);
        const uint32_t *ord = ordering + l * n;
        for (long k = 0; k < n; k++) out[k] += buf[ord[k]];
    }
    free(buf);
    return 0;
}

/* ---- ldpc/sparc_ldpc.py:72-77 + :128-134  Ay of block_sub_fht:
 * scatter y at ordering[l] into a zero w-vector, transform, keep the last m. */
int orc_block_Ay(const uint32_t *ordering, long L, long n, long m, long w,
                 const double *y, double *out)
{
    double *buf = (double *)malloc(sizeof(double) * (size_t)w);
    if (!buf) return -1;
    for (long l = 0; l < L; l++) {
        memset(buf, 0, sizeof(double) * (size_t)w);
        const uint32_t *ord = ordering + l * n;
        for (long k = 0; k < n; k++) buf[ord[k]] = y[k];
        orc_fht_inplace(buf, w);
        memcpy(out + l * m, buf + (w - m), sizeof(double) * (size_t)m);
    }
    free(buf);
    return 0;
}

/* ---- ldpc/src/c_ldpc.c:234-251  Lxor ------------------------------------- */
double orc_Lxor(double L1, double L2, int corr)
{
    double L = (signbit(L1) == signbit(L2)) ? 1.0 : -1.0;
    L *= fmin(fabs(L1), fabs(L2));
    if (corr) {
        L += log(1 + exp(-fabs(L1 + L2)));
        L -= log(1 + exp(-fabs(L1 - L2)));
    }
    return L;
}

/* ---- ldpc/src/c_ldpc.c:294-314  Lxfb: forward/backward extrinsics in place,
 * returns the total b[0]. ------------------------------------------------- */
#define ORC_MAXDC 64
double orc_Lxfb(double *L, long dc, int corr)
{
    double f[ORC_MAXDC], b[ORC_MAXDC];
    f[0] = L[0];
    b[dc - 1] = L[dc - 1];
    for (long k = 1; k < dc; k++) {
        f[k] = orc_Lxor(f[k - 1], L[k], corr);
        b[dc - k - 1] = orc_Lxor(b[dc - k], L[dc - k - 1], corr);
    }
    L[0] = b[1];
    L[dc - 1] = f[dc - 2];
    for (long k = 1; k < dc - 1; k++) L[k] = orc_Lxor(f[k - 1], b[k + 1], corr);
    return b[0];
}

static void orc_vn_step(const double *ch, const long *vdeg, const long *intrlv,
                        int Nv, double *msg, double *app)
{
    long p = 0;
    for (int j = 0; j < Nv; j++) {
        double aggr = ch[j];
        for (long k = 0; k < vdeg[j]; k++) aggr += msg[intrlv[p + k]];
        for (long k = 0; k < vdeg[j]; k++) msg[intrlv[p + k]] = aggr - msg[intrlv[p + k]];
        app[j] = aggr;
        p += vdeg[j];
    }
}

/* ---- ldpc/src/c_ldpc.c:138-206  sumprod2 --------------------------------- */
int orc_sumprod2(const double *ch, const long *vdeg, const long *cdeg, const long *intrlv,
                 int Nv, int Nc, int Nmsg, double *app, int max_it)
{
    double *msg = (double *)calloc((size_t)Nmsg, sizeof(double));
    if (!msg) return -1;
    int it;
    for (it = 0; it < max_it; it++) {
        orc_vn_step(ch, vdeg, intrlv, Nv, msg, app);
        int unsat = 0;
        long p = 0;
        for (int j = 0; j < Nc; j++) {
            double tot = orc_Lxfb(msg + p, cdeg[j], 1);
            if (tot <= 0.0) unsat = 1;
            p += cdeg[j];
        }
        if (!unsat) break;
    }
    free(msg);
    return it;
}

/* ---- ldpc/src/c_ldpc.c:32-113  sumprod (tanh / atanh rule) --------------- */
int orc_sumprod(const double *ch, const long *vdeg, const long *cdeg, const long *intrlv,
                int Nv, int Nc, int Nmsg, double *app, int max_it)
{
    double *msg = (double *)calloc((size_t)Nmsg, sizeof(double));
    if (!msg) return -1;
    int it;
    for (it = 0; it < max_it; it++) {
        orc_vn_step(ch, vdeg, intrlv, Nv, msg, app);
        int unsat = 0;
        long p = 0;
        for (int j = 0; j < Nc; j++) {
            double aggr = 1.0;
            for (long k = 0; k < cdeg[j]; k++) aggr *= (msg[p + k] = tanh(msg[p + k] / 2.0));
            /* the reference skips the atanh once a check is already unsatisfied
             * (:95); the value of the flag is the same either way */
            if (!unsat && 2.0 * atanh(aggr) <= 0.0) unsat = 1;
            for (long k = 0; k < cdeg[j]; k++) msg[p + k] = 2.0 * atanh(aggr / msg[p + k]);
            p += cdeg[j];
        }
        if (!unsat) break;
    }
    free(msg);
    return it;
}

/* ---- ldpc/src/c_ldpc.c:339-381  minsum, WITHOUT the reference's indexing bug
 * at :364 (`imsg += cdeg[j]` evaluated after `j++`, which misplaces every
 * check whose degree differs from its successor's and reads cdeg[Nc]).  For
 * check-regular codes (e.g. 802.16 rate 5/6) the two coincide except for the
 * out-of-bounds read. ------------------------------------------------------ */
int orc_minsum(const double *ch, const long *vdeg, const long *cdeg, const long *intrlv,
               int Nv, int Nc, int Nmsg, double *app, double corr_factor, int max_it)
{
    double *msg = (double *)calloc((size_t)Nmsg, sizeof(double));
    if (!msg) return -1;
    int it;
    for (it = 0; it < max_it; it++) {
        orc_vn_step(ch, vdeg, intrlv, Nv, msg, app);
        int unsat = 0;
        long p = 0;
        for (int j = 0; j < Nc; j++) {
            double tot = orc_Lxfb(msg + p, cdeg[j], 0);
            if (tot <= 0.0) unsat = 1;
            for (long k = 0; k < cdeg[j]; k++) msg[p + k] *= corr_factor;
            p += cdeg[j];
        }
        if (!unsat) break;
    }
    free(msg);
    return it;
}
