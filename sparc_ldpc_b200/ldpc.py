"""Host-side mirror of the reference's `ldpc.code` object (ldpc/py/ldpc.py:6-943).

Same attribute and method names (`proto, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, N, K, encode, decode, pcmat,
Lxor, Lxfb`), same NameError behaviour for bad parameters, but:
  * protograph tables come from data/protographs.json (extracted by tools/extract_protographs.py),
  * `prepare_decoder` uses closed-form quasi-cyclic edge addressing instead of the reference's
    port search (ldpc.py:694-786) -- identical arrays, milliseconds instead of 0.1 s,
  * `decode` calls libsparc_b200's `sumprod2/sumprod/minsum` symbols (the reference's own FFI,
    ldpc.py:923-927) which run on the GPU; `decode_batch` keeps everything on the device,
  * code objects are cached (the reference rebuilds one per codeword, sparc_ldpc.py:596).
"""
import ctypes as ct
import json
import os

import numpy as np

from . import _lib

MAX_ITCOUNT = 200
_DB = None
_HERE = os.path.dirname(os.path.abspath(__file__))


def _db():
    global _DB
    if _DB is None:
        with open(os.path.join(_HERE, "data", "protographs.json")) as f:
            _DB = json.load(f)
    return _DB


def _lookup_proto(standard, rate, z, ptype):
    db = _db()
    if standard == "802.11n":
        if z not in (27, 54, 81):
            raise NameError("802.11n invalid z (must be 27,54 or 81)")
        key, bad = "802.11n|%s|-|%d" % (rate, z), "802.11n invalid rate"
    elif standard == "802.16":
        if rate in ("2/3", "3/4"):
            if ptype not in ("A", "B"):
                raise NameError("802.16 type must be either A or B")
            key = "802.16|%s|%s|*" % (rate, ptype)
        else:
            key = "802.16|%s|-|*" % rate
        bad = "802.16 invalid rate"
    else:
        key, bad = "%s|%s|-|*" % (standard, rate), "IEEE standard unknown"
    if key not in db:
        raise NameError(bad)
    ent = db[key]
    proto = np.full((ent["rows"], ent["cols"]), -1, dtype=np.int64)
    e = np.asarray(ent["edges"], dtype=np.int64)
    proto[e[:, 0], e[:, 1]] = e[:, 2]
    return proto


class code:
    def __init__(self, standard="802.11n", rate="1/2", z=27, ptype="A"):
        self.standard, self.rate, self.z, self.ptype = standard, rate, z, ptype
        self.proto = self.assign_proto()
        self.vdeg, self.cdeg, self.intrlv = self.prepare_decoder()
        self.Nv, self.Nc, self.Nmsg = len(self.vdeg), len(self.cdeg), len(self.intrlv)
        self.N = self.Nv
        self.K = self.Nv - self.Nc
        self._graph = None
        self._enc = None

    def assign_proto(self):
        return _lookup_proto(self.standard, self.rate, self.z, self.ptype)

    # ------------------------------------------------------------------ decoder tables
    def prepare_decoder(self):
        """(vdeg, cdeg, intrlv) of ldpc.py:694-786 in closed form.

        Check r*z+k owns the consecutive check-major slots rowbase_r + k*dc_r + rank(r,c); the port of
        variable c*z+j that belongs to protograph row r (ports are filled in ascending r) reads slot
        rowbase_r + ((j - shift) mod z)*dc_r + rank(r,c), because check r*z+k meets variable
        c*z + (k+shift) mod z (ldpc.py:766-767).
        """
        proto, z = self.proto, int(self.z)
        nz = proto != -1
        dc, dv = nz.sum(1), nz.sum(0)
        cdeg = np.repeat(dc, z).astype(np.int64)
        vdeg = np.repeat(dv, z).astype(np.int64)
        rowbase = z * np.concatenate(([0], np.cumsum(dc)[:-1]))
        rank = np.cumsum(nz, axis=1) - 1
        j = np.arange(z)
        blocks = []
        for c in range(proto.shape[1]):
            rows = np.nonzero(nz[:, c])[0]
            if rows.size == 0:
                continue
            k = (j[:, None] - proto[rows, c][None, :]) % z          # [z, dv_c]
            blocks.append((rowbase[rows][None, :] + k * dc[rows][None, :] + rank[rows, c][None, :]).reshape(-1))
        intrlv = np.concatenate(blocks).astype(np.int64)
        return vdeg, cdeg, intrlv

    def pcmat(self):
        """Dense parity-check matrix (ldpc.py:666-691): H[r*z+k, c*z+(k+shift)%z] = 1."""
        proto, z = self.proto, int(self.z)
        H = np.zeros((z * proto.shape[0], z * proto.shape[1]), dtype=int)
        k = np.arange(z)
        for r, c in zip(*np.nonzero(proto != -1)):
            H[r * z + k, c * z + (k + proto[r, c]) % z] = 1
        return H

    # ------------------------------------------------------------------ encoder
    def _encoder_plan(self):
        if self._enc is None:
            proto, z = self.proto, int(self.z)
            Mp, Np = proto.shape
            Kp = Np - Mp
            hits = proto[:, Kp][proto[:, Kp] != -1] % z
            vals, cnt = np.unique(hits, return_counts=True)
            odd = vals[cnt % 2 == 1]
            if odd.size != 1:
                raise NameError("The offsets in colum Kp+1 of proto do not add to a single offset")
            self._enc = (Mp, Np, Kp, int(odd[0]))
        return self._enc

    def encode_batch(self, info):
        """Systematic QC encoding of a batch info[B, K] -> x[B, N] (ldpc.py:790-850), GF(2) on uint8."""
        z, proto = int(self.z), self.proto
        Mp, Np, Kp, toff = self._encoder_plan()
        info = np.asarray(info)
        if info.ndim != 2 or info.shape[1] != Kp * z:
            raise NameError("information word length not compatible with proto and z")
        B = info.shape[0]
        x = np.zeros((B, Np, z), dtype=np.uint8)
        x[:, :Kp] = (info.reshape(B, Kp, z) & 1).astype(np.uint8)
        col = np.arange(z)
        syn = np.zeros((B, Mp, z), dtype=np.uint8)           # systematic part of every block-row
        for r, c in zip(*np.nonzero(proto[:, :Kp] != -1)):
            syn[:, r] ^= x[:, c][:, (col + proto[r, c]) % z]
        total = np.bitwise_xor.reduce(syn, axis=1)
        x[:, Kp] = total[:, (col - toff) % z]                  # first parity block
        for r in range(Mp - 1):                                 # back-substitution down the dual diagonal
            acc = syn[:, r].copy()
            for c in np.nonzero(proto[r, Kp:Kp + r + 1] != -1)[0]:
                acc ^= x[:, Kp + c][:, (col + proto[r, Kp + c]) % z]
            x[:, Kp + r + 1] = acc
        return x.reshape(B, Np * z).astype(int)

    def encode(self, info):
        if len(info) != self.K:
            raise NameError("information word length not compatible with proto and z")
        return self.encode_batch(np.asarray(info).reshape(1, -1))[0]

    # ------------------------------------------------------------------ decoder
    def decode(self, ch, dectype="sumprod2", corr_factor=0.7):
        """Single codeword through the reference's FFI symbols (host pointers), computed on the GPU."""
        if len(ch) != len(self.vdeg):
            raise NameError("Channel inputs not consistent with variable degrees")
        L = _lib.lib()
        ch = np.ascontiguousarray(ch, dtype=np.double)
        app = np.zeros(self.Nv, dtype=np.double)
        args = (ch.ctypes.data, self.vdeg.ctypes.data, self.cdeg.ctypes.data, self.intrlv.ctypes.data,
                self.Nv, self.Nc, self.Nmsg, app.ctypes.data)
        if dectype == "sumprod":
            it = L.sumprod(*args)
        elif dectype == "sumprod2":
            it = L.sumprod2(*args)
        elif dectype == "minsum":
            it = L.minsum(*args, ct.c_double(corr_factor))
        else:
            raise NameError("Decoder type unknonwn")
        if it < 0:
            raise _lib.SparcB200Error("%s failed: %s" % (dectype, L.sb_last_error().decode()))
        return app, it

    def graph(self):
        """Device-resident Tanner graph handle for decode_batch (built once per code object)."""
        if self._graph is None:
            from .engine import Graph
            self._graph = Graph(self.vdeg, self.cdeg, self.intrlv)
        return self._graph

    def decode_batch(self, ch, dectype="sumprod2", corr_factor=0.7, max_it=MAX_ITCOUNT):
        """ch: CUDA float64 tensor [B, N] -> (app [B, N], it [B]) on the device."""
        return self.graph().bp(ch, dectype, corr_factor, max_it)

    def Lxor(self, L1, L2, corrflag=1):
        return _lib.lib().Lxor(L1, L2, corrflag)

    def Lxfb(self, L, corrflag=1):
        a = np.array(L, dtype=float)
        tot = _lib.lib().Lxfb(a.ctypes.data, len(a), corrflag)
        return tot, a


_CACHE = {}


def get_code(standard, rate, z, ptype="A"):
    """Cached `code` objects (the reference rebuilds the tables for every codeword)."""
    key = (standard, rate, z, ptype)
    if key not in _CACHE:
        _CACHE[key] = code(standard, rate, z, ptype)
    return _CACHE[key]
