"""Batched link decoders: the bodies of the reference's four per-codeword simulations
(ldpc/sparc_ldpc.py:359-1046) and of calc_E (ldpc/amp_exit.py:185-270) with the codeword loop turned into
the batch dimension.  Everything between "y is on the device" and "per-stage decisions / LLRs are on the
device" is libsparc_b200 kernels; the only torch ops are allocations, slices and copies.

Each flow returns a `Stages` object: per-stage section decisions idx [B, L] (int32, device) in the order
the reference reports its BER lists, plus iteration statistics for the roofline accounting.
"""
import numpy as np
import torch

from . import engine as E
from .ldpc import get_code

F64, I32 = torch.float64, torch.int32


class LinkSetup:
    """Derived sizes of a (SPARCParams, LDPCParams) pair (sparc_ldpc.py:370-416), device tables included."""

    def __init__(self, L, M, P, r, T, Pl=None, code=None, seed=0, device=None):
        self.L, self.M, self.P, self.r, self.T = int(L), int(M), P, r, int(T)
        self.n = int(L * np.log2(M) / r)
        self.logm = int(np.log2(M))
        self.total_bits = int(self.logm * L)
        self.Pl = np.asarray(P / L * np.ones(L) if Pl is None else Pl, dtype=np.float64)
        self.code = code
        self.nl = 0 if code is None else code.N
        self.kl = 0 if code is None else code.K
        if code is not None:
            assert self.nl <= self.total_bits
            # the LDPC must cover a whole number of sections (sparc_ldpc.py:416)
            assert self.nl % self.logm == 0
        self.ls = self.nl // self.logm
        self.R = (L * np.log2(M) - (self.nl - self.kl)) / self.n
        self.op = E.get_operator(self.L, self.M, self.n, seed)
        self.dev = E._dev() if device is None else device
        self.Pl_dev = torch.from_numpy(self.Pl).to(self.dev)
        self.graph = None if code is None else code.graph()

    def scale(self):
        """sqrt(n * repeat(Pl, M)) on the host (sparc_ldpc.py:657)."""
        return np.sqrt(self.n * np.repeat(self.Pl, self.M))


class Stages:
    def __init__(self):
        self.amp_idx, self.ldpc_idx = [], []      # decisions after each AMP / LDPC round
        self.amp_exec, self.bp_it = [], []        # executed AMP iterations [B] / BP iterations [B] per call
        self.amp_flags = []                       # SB_AMP_* status bits [B] per AMP call
        self.amp_sections = []                    # sections processed per AMP call ([B] tensor or int)
        self.extra = {}

    def ber(self, tx_idx, total_bits):
        """(ber_amp [B, n_amp], ber_ldpc [B, n_ldpc]) exactly as the reference divides (sparc_ldpc.py:650)."""
        def rows(lst):
            if not lst:
                return np.zeros((tx_idx.shape[0], 0))
            errs = torch.stack([E.count_errors(i, tx_idx) for i in lst], dim=1).cpu().numpy()
            return errs / total_bits
        return rows(self.amp_idx), rows(self.ldpc_idx)

    def ref_nan_count(self):
        """Codewords on which the reference's own global-max softmax would have left fp64's normal range in some
        AMP call (SB_AMP_REF_NAN): the kernel computes them accurately, the reference would not."""
        from ._lib import SB_AMP_REF_NAN
        if not self.amp_flags:
            return torch.zeros((), dtype=torch.int64)
        f = torch.stack([(x & SB_AMP_REF_NAN) != 0 for x in self.amp_flags]).any(dim=0)
        return f.sum()   # 0-dim device tensor: no host synchronisation here


def _amp(su, st, y, beta0=None, sections=None, nsec=None, Pl=None):
    res = su.op.amp(y, su.Pl_dev if Pl is None else Pl, su.T, beta0=beta0, sections=sections, nsec=nsec)
    st.amp_exec.append(res.n_exec)
    st.amp_flags.append(res.flags)   # SB_AMP_STOPPED / SB_AMP_REF_NAN per codeword (include/sparc_b200.h)
    st.amp_sections.append(su.L if nsec is None else nsec)
    return res


def plain(su, y):
    """AMP only (amp_ldpc_sim with ldpcparams=None, sparc_ldpc.py:448-462)."""
    st = Stages()
    res = _amp(su, st, y)
    st.amp_idx.append(E.argmax_sections(res.beta, su.L, su.M))
    st.extra["beta"] = res.beta
    return st


def _ldpc_round(su, st, beta, rx):
    """sp2bp -> LLR -> BP -> hard decisions on the protected sections (sparc_ldpc.py:657-681)."""
    L, M, ls = su.L, su.M, su.ls
    llr = E.sp2bp_llr(beta, M, su.n, su.Pl_dev, beta_first=L - ls, first_sec=L - ls, out_first=0, count=ls)
    app, it = su.graph.bp(llr)
    st.bp_it.append(it)
    idx = rx.clone()
    E.llr2idx(app, ls, M, out=idx[:, L - ls:])
    st.ldpc_idx.append(idx)
    return llr, app, idx


def soft(su, y, soft_iter):
    """soft_amp_ldpc_sim (sparc_ldpc.py:636-706): LDPC posteriors re-enter AMP as its initial beta."""
    st = Stages()
    res = _amp(su, st, y)
    beta = res.beta
    rx = E.argmax_sections(beta, su.L, su.M)
    st.amp_idx.append(rx)
    for _ in range(soft_iter):
        _, app, _ = _ldpc_round(su, st, beta, rx)
        beta0 = E.bp2sp_prior(app, su.ls, beta, su.L, su.M, su.n, su.Pl_dev, True)
        res = _amp(su, st, y, beta0=beta0)
        beta = res.beta
        rx = E.argmax_sections(beta, su.L, su.M)
        st.amp_idx.append(rx)
    st.extra["beta"] = beta
    return st


def hard_init(su, y):
    """hardinitbeta_amp_ldpc_sim (sparc_ldpc.py:795-854): hard-decided beta as the AMP initialisation."""
    st = Stages()
    res = _amp(su, st, y)
    rx = E.argmax_sections(res.beta, su.L, su.M)
    st.amp_idx.append(rx)
    _, _, idx = _ldpc_round(su, st, res.beta, rx)
    beta0 = E.onehot_beta(idx, su.Pl_dev, su.n, su.L, su.M)
    res = _amp(su, st, y, beta0=beta0)
    st.amp_idx.append(E.argmax_sections(res.beta, su.L, su.M))
    st.extra["beta"] = res.beta
    return st


def original_hard(su, y):
    """amp_ldpc_sim with an outer code (sparc_ldpc.py:448-539): peel the LDPC-decided sections off y and
    re-run AMP on the unprotected ones."""
    st = Stages()
    L, M, ls = su.L, su.M, su.ls
    res = _amp(su, st, y)
    rx = E.argmax_sections(res.beta, L, M)
    st.amp_idx.append(rx)
    _, _, idx = _ldpc_round(su, st, res.beta, rx)
    Lu = L - ls
    if Lu > 0:
        B = y.shape[0]
        peel = idx.clone()
        peel[:, :Lu] = -1
        y_new = su.op.onehot_apply(peel, su.Pl_dev, y, sign=-1.0)
        sections = torch.arange(L, dtype=I32, device=y.device).repeat(B, 1).contiguous()
        nsec = torch.full((B,), Lu, dtype=I32, device=y.device)
        res2 = _amp(su, st, y_new, sections=sections, nsec=nsec)
        final = idx.clone()
        E.argmax_sections(res2.beta, Lu, M, out=final[:, :Lu])
        st.amp_idx.append(final)
    return st


def threshold(su, y, soft_iter, thr):
    """soft_amp_ldpc_hardinit (sparc_ldpc.py:953-1040): threshold-peeled sections are hard decided, AMP restarts
    from zero on the rest, LLRs of peeled sections keep their pre-AMP (LDPC) values."""
    st = Stages()
    L, M, ls, logm = su.L, su.M, su.ls, su.logm
    res = _amp(su, st, y)
    st.amp_idx.append(E.argmax_sections(res.beta, L, M))
    LLR = E.sp2bp_llr(res.beta, M, su.n, su.Pl_dev, count=L)  # all sections (sparc_ldpc.py:977-981)
    beta_unprot = res.beta
    peeled = []
    for i in range(soft_iter):
        ch = LLR[:, (L - ls) * logm:].contiguous()
        app, it = su.graph.bp(ch)
        st.bp_it.append(it)
        LLR[:, (L - ls) * logm:] = app
        st.ldpc_idx.append(E.llr2idx(LLR, L, M))
        if i == soft_iter - 1:
            break
        post = E.bp2sp_prior(app, ls, beta_unprot, L, M, su.n, su.Pl_dev, False)
        hard, act, nact = E.threshold_peel(post, L, M, ls, thr)
        peeled.append(nact)
        y_new = su.op.onehot_apply(hard, su.Pl_dev, y, sign=-1.0)
        res = _amp(su, st, y_new, sections=act, nsec=nact)
        E.sp2bp_llr(res.beta, M, su.n, su.Pl_dev, sections=act, nsec=nact, out=LLR)
        st.amp_idx.append(E.llr2idx(LLR, L, M))
    st.extra["LLR"] = LLR
    st.extra["nact"] = peeled
    return st


def exit_E(su, y, A, thr):
    """calc_E (amp_exit.py:213-259): a-priori LLRs A [B, L*logM] -> extrinsic LLRs E (clipped to +-55)."""
    st = Stages()
    L, M = su.L, su.M
    post = E.bp2sp_prior(A, L, None, L, M, su.n, su.Pl_dev, False)     # bp2sp(1/(1+exp(A)))  (:223-225)
    hard, act, nact = E.threshold_peel(post, L, M, L, thr)             # ldpc_sections=None -> all (:75-77)
    y_new = su.op.onehot_apply(hard, su.Pl_dev, y, sign=-1.0)
    res = _amp(su, st, y_new, sections=act, nsec=nact)
    Eo = A.clone()
    E.sp2bp_llr(res.beta, M, su.n, su.Pl_dev, sections=act, nsec=nact, out=Eo)
    Eo.clamp_(-55, 55)                                                  # np.clip (:259)
    st.extra["nact"] = nact
    return Eo, st


def make_setup(sparcparams, ldpcparams=None, seed=0):
    """LinkSetup from reference-style parameter objects (sparc_ldpc.py:227-255)."""
    from .sparc_ldpc import pa_parameterised
    sp = sparcparams
    Pl = None
    if sp.a is not None:
        Pl = pa_parameterised(sp.L, sp.C, sp.p, sp.a, sp.f)
    code = None
    if ldpcparams is not None:
        code = get_code(ldpcparams.standard, ldpcparams.r_ldpc, ldpcparams.z, ldpcparams.ptype)
    return LinkSetup(sp.L, sp.M, sp.p, sp.r, sp.t, Pl=Pl, code=code, seed=seed)
