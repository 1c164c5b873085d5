"""Device-side engine: thin torch wrappers over the C ABI (include/sparc_b200.h).

torch is used for what it is good at here -- device memory, streams, torch.distributed -- and nothing
else: every computation below is one call into libsparc_b200.so with raw device pointers.  All tensors
are CUDA float64 / int32, row-major, codeword-major ([B, ...]).
"""
import numpy as np
import torch

from . import _lib
from ._lib import check

F64, I32 = torch.float64, torch.int32
_RULES = {"sumprod2": _lib.SB_BP_SUMPROD2, "sumprod": _lib.SB_BP_SUMPROD, "minsum": _lib.SB_BP_MINSUM,
          "sumprod2_fast": _lib.SB_BP_SUMPROD2_FAST}
_MODES = {"strict": _lib.SB_AMP_STRICT, "fast": _lib.SB_AMP_FAST, "f64": _lib.SB_AMP_F64}
# Arithmetic of Operator.amp.  "f64" (default) = fp64 throughout and the reference's exact-equality stop rule; at
# M = 512 the gathers add their terms in a bank-scheduled order (warp-specialised kernel, 1.8x faster per iteration than
# "strict", from which it differs by fp64 summation-order noise, ~1e-15 per iteration), for every other shape it IS
# "strict".  "strict" = fp64 in the reference's order of additions everywhere.  "fast" = 32-bit fixed-point gathers and
# a tolerance stop rule.  Override per call with mode=..., or globally with the environment
# variable SPARC_B200_AMP_MODE.
import ctypes as _ct
import os as _os
AMP_MODE = _os.environ.get("SPARC_B200_AMP_MODE", "f64")
# Default arithmetic of Graph.bp for dectype "sumprod2": "strict" = fp64 exp/log as c_ldpc.c:246-247; "fast" = the
# Lxor correction terms in single precision (SB_BP_SUMPROD2_FAST).  Environment: SPARC_B200_BP_MODE.
BP_MODE = _os.environ.get("SPARC_B200_BP_MODE", "strict")


def _dev():
    if not torch.cuda.is_available():
        raise _lib.SparcB200Error("no CUDA device: libsparc_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA %s tensor" % (name, dtype))


def transform_width(M, n):
    """w of sparc_ldpc.py:54,110."""
    return 2 ** int(np.ceil(np.log2(max(M + 1, n + 1))))


def make_ordering(L, M, n, seed=0):
    """Row-selection table of the block sub-sampled Hadamard design (sparc_ldpc.py:110-117): cumulative
    in-place shuffles of arange(1, w) with RandomState(seed) -- host logic, shared by the whole batch."""
    w = transform_width(M, n)
    rng = np.random.RandomState(seed)
    ordering = np.empty((L, n), dtype=np.uint32)
    idxs = np.arange(1, w, dtype=np.uint32)
    for ll in range(L):
        rng.shuffle(idxs)
        ordering[ll] = idxs[:n]
    return ordering


class AmpResult:
    __slots__ = ("beta", "iters", "n_exec", "flags", "tau2")

    def __init__(self, beta, iters, n_exec, flags, tau2):
        self.beta, self.iters, self.n_exec, self.flags, self.tau2 = beta, iters, n_exec, flags, tau2


class Operator:
    """Device handle of the design operator for one (L, M, n, ordering)."""

    def __init__(self, L, M, n, ordering=None, seed=0):
        self.L, self.M, self.n = int(L), int(M), int(n)
        self.logM = int(np.log2(M))
        if ordering is None:
            ordering = make_ordering(L, M, n, seed)
        self.ordering = np.ascontiguousarray(ordering, dtype=np.uint32)
        if self.ordering.shape != (self.L, self.n):
            raise ValueError("ordering must have shape (L, n)")
        _dev()
        import ctypes as ct
        h = ct.c_void_p()
        check(_lib.lib().sb_operator_create(self.ordering.ctypes.data, self.L, self.M, self.n, ct.byref(h)),
              "sb_operator_create")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sb_operator_destroy(h)
            except Exception:
                pass
            self._h = None

    def _lists(self, sections, nsec, B):
        if sections is None:
            return None, None
        _chk(sections, I32, "sections")
        _chk(nsec, I32, "nsec")
        if sections.shape != (B, self.L) or nsec.shape != (B,):
            raise ValueError("sections must be [B, L] and nsec [B]")
        return sections, nsec

    def Ab(self, beta, sections=None, nsec=None):
        """[B, L*M] (compact over the section list) -> A beta [B, n]   (sparc_ldpc.py:143-144)."""
        _chk(beta, F64, "beta")
        B = beta.shape[0]
        if beta.shape[1] != self.L * self.M:
            raise ValueError("beta rows must have L*M entries")
        sections, nsec = self._lists(sections, nsec, B)
        out = torch.empty((B, self.n), dtype=F64, device=beta.device)
        check(_lib.lib().sb_Ab_batch(self._h, _p(beta), _p(sections), _p(nsec), B, _p(out), _stream()), "sb_Ab_batch")
        return out

    def Az(self, z, sections=None, nsec=None):
        """[B, n] -> A^T z [B, L*M]   (sparc_ldpc.py:145-146); rows beyond nsec*M are zero."""
        _chk(z, F64, "z")
        B = z.shape[0]
        sections, nsec = self._lists(sections, nsec, B)
        out = torch.zeros((B, self.L * self.M), dtype=F64, device=z.device)
        check(_lib.lib().sb_Az_batch(self._h, _p(z), _p(sections), _p(nsec), B, _p(out), _stream()), "sb_Az_batch")
        return out

    def onehot_apply(self, idx, Pl, y=None, sign=1.0):
        """y + sign * A beta_onehot(idx) with beta_onehot[l*M+idx[l]] = sqrt(n Pl[l]); idx < 0 skips a section."""
        _chk(idx, I32, "idx")
        _chk(Pl, F64, "Pl")
        _chk(y, F64, "y")
        B = idx.shape[0]
        out = torch.empty((B, self.n), dtype=F64, device=idx.device)
        check(_lib.lib().sb_onehot_apply_batch(self._h, _p(idx), _p(Pl), _p(y), float(sign), B, _p(out), _stream()),
              "sb_onehot_apply_batch")
        return out

    def amp(self, y, Pl, T, beta0=None, sections=None, nsec=None, trace=False, mode=None):
        """Batched AMP decode (sparc_ldpc.py:189-222).  Returns AmpResult.  mode: "strict" | "fast" (default:
        engine.AMP_MODE, see include/sparc_b200.h SB_AMP_STRICT / SB_AMP_FAST)."""
        mode = AMP_MODE if mode is None else mode
        if mode not in _MODES:
            raise ValueError("mode must be 'strict', 'f64' or 'fast'")
        _chk(y, F64, "y")
        _chk(Pl, F64, "Pl")
        _chk(beta0, F64, "beta0")
        B = y.shape[0]
        if y.shape[1] != self.n or Pl.numel() != self.L:
            raise ValueError("y must be [B, n] and Pl [L]")
        sections, nsec = self._lists(sections, nsec, B)
        dev = y.device
        alloc = torch.empty if sections is None else torch.zeros
        beta = alloc((B, self.L * self.M), dtype=F64, device=dev)
        iters = torch.empty(B, dtype=I32, device=dev)
        n_exec = torch.empty(B, dtype=I32, device=dev)
        flags = torch.empty(B, dtype=I32, device=dev)
        tau2 = torch.full((B, max(T, 1)), float("nan"), dtype=F64, device=dev) if trace else None
        scratch = torch.empty((2, B, self.n), dtype=F64, device=dev) if mode in ("fast", "f64") else None
        check(_lib.lib().sb_amp_batch(self._h, _p(y), _p(Pl), _p(beta0), _p(sections), _p(nsec), B, int(T), _MODES[mode],
                                      _p(beta), _p(iters), _p(n_exec), _p(flags), _p(tau2), _p(scratch), _stream()),
              "sb_amp_batch")
        return AmpResult(beta, iters, n_exec, flags, tau2)


def section_softmax(s, Pl, tau2, L, M, n):
    """Denoiser of amp() alone (sparc_ldpc.py:214-219) for a batch: s [B, L*M] = beta + A^T z, tau2 [B] ->
    (beta [B, L*M], sumsq [B, L]).  For design operators that are not ours (foreign Ab / Az closures)."""
    _chk(s, F64, "s")
    _chk(Pl, F64, "Pl")
    _chk(tau2, F64, "tau2")
    B = s.shape[0]
    if s.shape[1] != L * M or Pl.numel() != L or tau2.numel() != B:
        raise ValueError("s must be [B, L*M], Pl [L], tau2 [B]")
    beta = torch.empty_like(s)
    sumsq = torch.empty((B, L), dtype=F64, device=s.device)
    check(_lib.lib().sb_section_softmax_batch(_p(s), _p(Pl), _p(tau2), 0, int(L), int(M), int(n), B, _p(beta), _p(sumsq),
                                              _stream()), "sb_section_softmax_batch")
    return beta, sumsq


class DenseOperator:
    """Dense design matrix A [n, L*M] on the device (Gaussian-A mode; the reference's amp() accepts any Ab/Az
    closures, sparc_ldpc.py:189).  A beta and A^T z are libsparc_b200's tcgen05 / TMA GEMMs over the batch with a
    bf16x3 split of every operand (FP32 emulation, csrc/dense.cu); the AMP loop, the section softmax and the
    Onsager / tau^2 reductions are fp64 kernels of the same library.  Results agree with an fp64 evaluation of
    the same A to ~1e-7 (tolerance of the north star: 1e-5)."""

    def __init__(self, A, L, M):
        import ctypes as ct
        dev = _dev()
        A = torch.as_tensor(A, dtype=F64).to(dev).contiguous()
        self.n, LM = A.shape
        self.L, self.M = int(L), int(M)
        if LM != self.L * self.M:
            raise ValueError("A must have L*M columns")
        h = ct.c_void_p()
        check(_lib.lib().sb_dense_create(A.data_ptr(), int(self.n), int(LM), ct.byref(h)), "sb_dense_create")
        torch.cuda.synchronize()
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sb_dense_destroy(h)
            except Exception:
                pass
            self._h = None

    def _apply(self, x, transpose):
        _chk(x, F64, "x")
        B, K = x.shape
        want = self.n if transpose else self.L * self.M
        if K != want:
            raise ValueError("operand rows must have %d entries" % want)
        out = torch.empty((B, self.L * self.M if transpose else self.n), dtype=F64, device=x.device)
        check(_lib.lib().sb_dense_apply_batch(self._h, int(transpose), _p(x), B, _p(out), _stream()), "sb_dense_apply_batch")
        return out

    def Ab(self, beta):
        """[B, L*M] -> A beta [B, n]"""
        return self._apply(beta, 0)

    def Az(self, z):
        """[B, n] -> A^T z [B, L*M]"""
        return self._apply(z, 1)

    def amp(self, y, Pl, T, beta0=None, trace=False):
        """sparc_ldpc.py:189-222 for a batch; returns AmpResult (iters = amp_test's t)."""
        _chk(y, F64, "y")
        _chk(Pl, F64, "Pl")
        _chk(beta0, F64, "beta0")
        B, n = y.shape
        if n != self.n or Pl.numel() != self.L:
            raise ValueError("y must be [B, n] and Pl [L]")
        dev = y.device
        beta = torch.empty((B, self.L * self.M), dtype=F64, device=dev)
        iters = torch.empty(B, dtype=I32, device=dev)
        n_exec = torch.empty(B, dtype=I32, device=dev)
        flags = torch.empty(B, dtype=I32, device=dev)
        tau2 = torch.empty((B, max(T, 1)), dtype=F64, device=dev) if trace else None
        check(_lib.lib().sb_dense_amp_batch(self._h, _p(y), _p(Pl), _p(beta0), self.L, self.M, B, int(T), _p(beta),
                                            _p(iters), _p(n_exec), _p(flags), _p(tau2), _stream()), "sb_dense_amp_batch")
        return AmpResult(beta, iters, n_exec, flags, tau2)


    def amp_sharded(self, y, Pl_local, P_total, T, allreduce=None, group=None, beta0=None, trace=False):
        """Column-sharded AMP: this operator holds the columns of the local sections of a larger matrix
        (sb_dense_amp_batch_sharded).  y [B, n] is replicated; Pl_local / beta0 / the returned beta refer to the
        local sections; P_total = sum of Pl over all sections.  `allreduce(tensor)` must sum a CUDA float64 tensor
        in place over the ranks (default: torch.distributed.all_reduce over `group`, i.e. NCCL over NVLink); it is
        called once per AMP iteration on [B*n + B] doubles (partial A beta and |beta|^2)."""
        import ctypes as ct
        _chk(y, F64, "y")
        _chk(Pl_local, F64, "Pl_local")
        _chk(beta0, F64, "beta0")
        B, n = y.shape
        if n != self.n or Pl_local.numel() != self.L:
            raise ValueError("y must be [B, n] and Pl_local [L_local]")
        dev = y.device
        if allreduce is None:
            import torch.distributed as dist

            def allreduce(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        xbuf = torch.empty(B * n + B, dtype=F64, device=dev)
        err = []

        def _cb(ctx, buf, count, stream):
            try:
                allreduce(xbuf)
                return 0
            except Exception as ex:  # never let an exception cross the C ABI
                err.append(ex)
                return 1

        cb = ct.CFUNCTYPE(ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_long, ct.c_void_p)(_cb)
        beta = torch.empty((B, self.L * self.M), dtype=F64, device=dev)
        iters = torch.empty(B, dtype=I32, device=dev)
        n_exec = torch.empty(B, dtype=I32, device=dev)
        flags = torch.empty(B, dtype=I32, device=dev)
        tau2 = torch.empty((B, max(T, 1)), dtype=F64, device=dev) if trace else None
        rc = _lib.lib().sb_dense_amp_batch_sharded(self._h, _p(y), _p(Pl_local), float(P_total), _p(beta0), self.L, self.M,
                                                   B, int(T), _p(beta), _p(iters), _p(n_exec), _p(flags), _p(tau2),
                                                   _p(xbuf), ct.cast(cb, ct.c_void_p), None, _stream())
        if err:
            raise err[0]
        check(rc, "sb_dense_amp_batch_sharded")
        return AmpResult(beta, iters, n_exec, flags, tau2)


    def amp_p2p(self, y, Pl_local, P_total, T, peers, beta0=None, trace=False, before_launch=None):
        """Column-sharded AMP with the per-iteration exchange over NVLink peer memory (sb_dense_amp_batch_p2p): no
        collective library in the loop.  `peers` is this rank's PeerExchange (receive areas and flag words of all
        ranks, mapped here).  Same arguments and result as amp_sharded.  `before_launch()` (optional) runs after
        this call's device allocations and right before the kernels are queued: ranks that are THREADS of one
        process rendezvous there, because a cudaMalloc of one thread can wait for the whole device, i.e. for a peer
        thread's kernel that is already spinning on this thread's flag (separate processes need nothing)."""
        import ctypes as ct
        _chk(y, F64, "y")
        _chk(Pl_local, F64, "Pl_local")
        _chk(beta0, F64, "beta0")
        B, n = y.shape
        if n != self.n or Pl_local.numel() != self.L:
            raise ValueError("y must be [B, n] and Pl_local [L_local]")
        if B * n + B > peers.slot_doubles:
            raise ValueError("the peer areas were allocated for a smaller batch")
        dev = y.device
        xbuf = torch.empty(B * n + B, dtype=F64, device=dev)
        beta = torch.empty((B, self.L * self.M), dtype=F64, device=dev)
        iters = torch.empty(B, dtype=I32, device=dev)
        n_exec = torch.empty(B, dtype=I32, device=dev)
        flags = torch.empty(B, dtype=I32, device=dev)
        tau2 = torch.empty((B, max(T, 1)), dtype=F64, device=dev) if trace else None
        if before_launch is not None:
            before_launch()
        rc = _lib.lib().sb_dense_amp_batch_p2p(self._h, _p(y), _p(Pl_local), float(P_total), _p(beta0), self.L, self.M, B,
                                               int(T), _p(beta), _p(iters), _p(n_exec), _p(flags), _p(tau2), _p(xbuf),
                                               ct.addressof(peers.cstruct), ct.addressof(peers.epoch), _stream())
        check(rc, "sb_dense_amp_batch_p2p")
        return AmpResult(beta, iters, n_exec, flags, tau2)


class _SbP2p(_ct.Structure):
    """struct sb_p2p of include/sparc_b200.h"""
    _fields_ = [("rank", _ct.c_int), ("world", _ct.c_int), ("timeout_ms", _ct.c_int),
                ("slots", _ct.c_void_p * 8), ("flags", _ct.c_void_p * 8), ("slot_doubles", _ct.c_long)]


class PeerExchange:
    """Receive areas ([2][world][B n + B] doubles) and flag words ([world] uint64) of every rank of a column-sharded
    dense operator, mapped into this process (include/sparc_b200.h sb_p2p)."""

    def __init__(self, rank, world, slot_doubles, areas, flagws, timeout_ms=20000):
        import ctypes as ct
        if world > 8:
            raise ValueError("at most 8 ranks (one NVSwitch domain)")
        self.rank, self.world, self.slot_doubles = rank, world, slot_doubles
        self._keep = (areas, flagws)          # the mapped tensors must outlive the exchange
        self.cstruct = _SbP2p()
        self.cstruct.rank, self.cstruct.world, self.cstruct.timeout_ms = rank, world, int(timeout_ms)
        self.cstruct.slot_doubles = int(slot_doubles)
        for r in range(world):
            self.cstruct.slots[r] = areas[r].data_ptr()
            self.cstruct.flags[r] = flagws[r].data_ptr()
        self.epoch = ct.c_ulonglong(0)

    @staticmethod
    def _alloc(world, B, n, dev):
        S = B * n + B
        return S, torch.zeros(2 * world * S, dtype=F64, device=dev), torch.zeros(world, dtype=torch.int64, device=dev)

    @staticmethod
    def in_process(world, B, n, dev=None):
        """All ranks live in this process (threads, one stream each).  Only safe when the ranks' streams never share a
        hardware queue and no thread allocates device memory while a peer spins on its flag (use amp_p2p's
        before_launch rendezvous); separate processes (from_process_group) have neither restriction."""
        dev = dev or _dev()
        allocs = [PeerExchange._alloc(world, B, n, dev) for _ in range(world)]
        areas, flagws = [a[1] for a in allocs], [a[2] for a in allocs]
        torch.cuda.synchronize()
        return [PeerExchange(r, world, allocs[0][0], areas, flagws) for r in range(world)]

    @staticmethod
    def from_process_group(B, n, group=None, timeout_ms=20000):
        """One process per GPU (torchrun): every rank allocates its area (sb_p2p_alloc: cudaMalloc, zeroed) on its own
        GPU, the 64-byte CUDA IPC handles travel through torch.distributed, and every rank opens its peers' handles
        with its own device current (sb_p2p_open), which maps the peers' memory for this GPU over NVLink."""
        import ctypes as ct
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        _dev()
        L = _lib.lib()
        S = B * n + B
        area_bytes = 2 * world * S * 8
        base, hbuf = ct.c_void_p(), (ct.c_ubyte * 64)()
        check(L.sb_p2p_alloc(area_bytes + world * 8, ct.byref(base), ct.addressof(hbuf)), "sb_p2p_alloc")
        everyone = [None] * world
        dist.all_gather_object(everyone, (torch.cuda.current_device(), bytes(hbuf)), group=group)
        bases = []
        for r, (pdev, handle) in enumerate(everyone):
            if r == rank:
                bases.append(base.value)
                continue
            if pdev != torch.cuda.current_device():
                check(L.sb_enable_peer_access(int(pdev)), "sb_enable_peer_access")
            p, hb = ct.c_void_p(), (ct.c_ubyte * 64).from_buffer_copy(handle)
            check(L.sb_p2p_open(ct.addressof(hb), ct.byref(p)), "sb_p2p_open")
            bases.append(p.value)
        dist.barrier(group=group)             # every area is zeroed and mapped before anyone pushes

        class _Ptr:                            # what PeerExchange needs from a tensor
            def __init__(self, addr):
                self.addr = addr

            def data_ptr(self):
                return self.addr

        px = PeerExchange(rank, world, S, [_Ptr(b) for b in bases], [_Ptr(b + area_bytes) for b in bases], timeout_ms)
        px._bases, px._group = bases, group
        return px

    def close(self):
        """Unmap the peers' areas and free the local one (collective: every rank calls it)."""
        bases = getattr(self, "_bases", None)
        if not bases:
            return
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self._group)       # nobody still pushes into memory that is about to go away
        for r, b in enumerate(bases):
            if r != self.rank:
                _lib.lib().sb_p2p_close(b)
        dist.barrier(group=self._group)
        _lib.lib().sb_p2p_free(bases[self.rank])
        self._bases = None


_OP_CACHE = {}


def get_operator(L, M, n, seed=0):
    """Operators with a fixed seed are shared by every codeword (sparc_ldpc.py:140 default seed=0)."""
    if seed is None:
        return Operator(L, M, n, seed=None)
    key = (L, M, n, seed, torch.cuda.current_device())
    if key not in _OP_CACHE:
        _OP_CACHE[key] = Operator(L, M, n, seed=seed)
    return _OP_CACHE[key]


class Graph:
    """Device handle of a Tanner graph given by (vdeg, cdeg, intrlv) of ldpc.py:694-786."""

    def __init__(self, vdeg, cdeg, intrlv):
        import ctypes as ct
        self.vdeg = np.ascontiguousarray(vdeg, dtype=np.int64)
        self.cdeg = np.ascontiguousarray(cdeg, dtype=np.int64)
        self.intrlv = np.ascontiguousarray(intrlv, dtype=np.int64)
        self.Nv, self.Nc, self.Nmsg = len(self.vdeg), len(self.cdeg), len(self.intrlv)
        _dev()
        h = ct.c_void_p()
        check(_lib.lib().sb_graph_create(self.vdeg.ctypes.data, self.cdeg.ctypes.data, self.intrlv.ctypes.data,
                                         self.Nv, self.Nc, self.Nmsg, ct.byref(h)), "sb_graph_create")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().sb_graph_destroy(h)
            except Exception:
                pass
            self._h = None

    def bp(self, ch, dectype="sumprod2", corr_factor=0.7, max_it=_lib.SB_MAX_ITCOUNT):
        _chk(ch, F64, "ch")
        if dectype not in _RULES:
            raise NameError("Decoder type unknonwn")
        B = ch.shape[0]
        if ch.shape[1] != self.Nv:
            raise NameError("Channel inputs not consistent with variable degrees")
        app = torch.empty_like(ch)
        it = torch.empty(B, dtype=I32, device=ch.device)
        rule = _RULES["sumprod2_fast"] if (dectype == "sumprod2" and BP_MODE == "fast") else _RULES[dectype]
        check(_lib.lib().sb_bp_batch(self._h, rule, _p(ch), B, _p(app), _p(it), int(max_it),
                                     float(corr_factor), _stream()), "sb_bp_batch")
        return app, it


# ---------------------------------------------------------------------------------------- handoff kernels
def sp2bp_llr(beta, M, n, Pl, beta_first=0, first_sec=0, out_first=0, count=None, sections=None, nsec=None,
              out=None, out_bits=None, want_p=False):
    """LLRs (and optionally bit posteriors) of sections of beta [B, *]; see sb_sp2bp_llr_batch."""
    _chk(beta, F64, "beta")
    _chk(Pl, F64, "Pl")
    B = beta.shape[0]
    logM = int(np.log2(M))
    Lstride = 0
    if sections is not None:
        _chk(sections, I32, "sections")
        _chk(nsec, I32, "nsec")
        Lstride = sections.shape[1]
        count = 0
    if out is None:
        if out_bits is None:
            out_bits = (out_first + count) * logM
        out = torch.zeros((B, out_bits), dtype=F64, device=beta.device)
    _chk(out, F64, "out")
    p = torch.zeros_like(out) if want_p else None
    check(_lib.lib().sb_sp2bp_llr_batch(_p(beta), beta.stride(0), int(beta_first), _p(sections), _p(nsec), int(Lstride),
                                        int(first_sec), int(out_first), int(count), int(M), int(n), _p(Pl), B, _p(p),
                                        _p(out), out.stride(0), _stream()), "sb_sp2bp_llr_batch")
    return (out, p) if want_p else out


def bp2sp_prior(app, ls, beta_prev, L, M, n, Pl, scale_by_power=True, from_prob=False):
    """Next AMP initialisation / section posterior from LDPC a-posteriori LLRs (sparc_ldpc.py:685-696)."""
    _chk(app, F64, "app")
    _chk(beta_prev, F64, "beta_prev")
    _chk(Pl, F64, "Pl")
    B = app.shape[0] if app is not None else beta_prev.shape[0]
    dev = app.device if app is not None else beta_prev.device
    out = torch.empty((B, L * M), dtype=F64, device=dev)
    check(_lib.lib().sb_bp2sp_prior_batch(_p(app), int(ls), _p(beta_prev), int(L), int(M), int(n), _p(Pl),
                                          (1 if scale_by_power else 0) | (2 if from_prob else 0), B, _p(out),
                                          _stream()), "sb_bp2sp_prior_batch")
    return out


def argmax_sections(beta, count, M, out=None):
    _chk(beta, F64, "beta")
    B = beta.shape[0]
    if out is None:
        out = torch.empty((B, count), dtype=I32, device=beta.device)
    check(_lib.lib().sb_argmax_batch(_p(beta), beta.stride(0), int(count), int(M), B, _p(out), out.stride(0), _stream()),
          "sb_argmax_batch")
    return out


def llr2idx(llr, count, M, out=None):
    """Hard decisions (llr < 0) packed MSB first into section indices; `out` may be a column slice view."""
    _chk_strided = llr.is_cuda and llr.dtype == F64 and llr.stride(1) == 1
    if not _chk_strided:
        raise ValueError("llr must be a CUDA float64 tensor with unit inner stride")
    B = llr.shape[0]
    if out is None:
        out = torch.empty((B, count), dtype=I32, device=llr.device)
    check(_lib.lib().sb_llr2idx_batch(llr.data_ptr(), llr.stride(0), int(count), int(M), B, out.data_ptr(),
                                      out.stride(0), _stream()), "sb_llr2idx_batch")
    return out


def count_errors(a, t):
    """Bit errors per codeword: sum_i popcount(a[b,i] ^ t[b,i]) (sparc_ldpc.py:650)."""
    _chk(a, I32, "a")
    _chk(t, I32, "t")
    B, count = a.shape
    out = torch.empty(B, dtype=I32, device=a.device)
    check(_lib.lib().sb_count_errors_batch(_p(a), _p(t), int(count), B, _p(out), _stream()), "sb_count_errors_batch")
    return out


def onehot_beta(idx, Pl, n, L, M):
    """Hard-decided beta (sparc_ldpc.py:840-843): sqrt(n Pl[l]) at idx[b, l], zero elsewhere."""
    _chk(idx, I32, "idx")
    _chk(Pl, F64, "Pl")
    B = idx.shape[0]
    beta = torch.empty((B, L * M), dtype=F64, device=idx.device)
    check(_lib.lib().sb_onehot_beta_batch(_p(idx), _p(Pl), int(n), int(L), int(M), B, _p(beta), _stream()),
          "sb_onehot_beta_batch")
    return beta


def ldpc_encode(code, info):
    """Device QC-LDPC encoder: info [B, K] uint8 (CUDA) -> codewords [B, N] uint8 (ldpc.py:790-850)."""
    if not (info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous()):
        raise ValueError("info must be a contiguous CUDA uint8 tensor")
    Mp, Np, Kp, toff = code._encoder_plan()
    z = int(code.z)
    if info.shape[1] != Kp * z:
        raise NameError("information word length not compatible with proto and z")
    if getattr(code, "_proto_dev", None) is None or code._proto_dev.device != info.device:
        code._proto_dev = torch.from_numpy(np.ascontiguousarray(code.proto, dtype=np.int32)).to(info.device)
    B = info.shape[0]
    x = torch.empty((B, Np * z), dtype=torch.uint8, device=info.device)
    check(_lib.lib().sb_ldpc_encode_batch(_p(code._proto_dev), Mp, Np, z, toff, _p(info), B, _p(x), _stream()),
          "sb_ldpc_encode_batch")
    return x


def bits2idx(bits, count, M, out=None):
    """MSB-first section indices of a CUDA uint8 bit tensor [B, >= count*logM] (sparc_ldpc.py:317-341)."""
    if not (bits.is_cuda and bits.dtype == torch.uint8 and bits.stride(1) == 1):
        raise ValueError("bits must be a CUDA uint8 tensor with unit inner stride")
    B = bits.shape[0]
    if out is None:
        out = torch.empty((B, count), dtype=I32, device=bits.device)
    check(_lib.lib().sb_bits2idx_batch(bits.data_ptr(), bits.stride(0), int(count), int(M), B, out.data_ptr(),
                                       out.stride(0), _stream()), "sb_bits2idx_batch")
    return out


def threshold_peel(post, L, M, ls, threshold):
    """amp_exit.py:85-106 -> (hard_idx [B, L], act [B, L], nact [B])."""
    _chk(post, F64, "post")
    B = post.shape[0]
    dev = post.device
    hard = torch.empty((B, L), dtype=I32, device=dev)
    act = torch.empty((B, L), dtype=I32, device=dev)
    nact = torch.empty(B, dtype=I32, device=dev)
    check(_lib.lib().sb_threshold_peel_batch(_p(post), int(L), int(M), int(ls), float(threshold), B, _p(hard), _p(act),
                                             _p(nact), _stream()), "sb_threshold_peel_batch")
    return hard, act, nact


def exit_hist(E, X, edges):
    """counts [B, 2, nbins] (row 0: X == +1, row 1: X == -1) with numpy.histogram semantics."""
    _chk(E, F64, "E")
    _chk(X, I32, "X")
    _chk(edges, F64, "edges")
    B, length = E.shape
    nb = edges.numel() - 1
    counts = torch.empty((B, 2, nb), dtype=torch.int64, device=E.device)
    check(_lib.lib().sb_exit_hist_batch(_p(E), _p(X), int(length), _p(edges), int(edges.numel()), B, _p(counts),
                                        _stream()), "sb_exit_hist_batch")
    return counts
