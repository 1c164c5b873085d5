"""sparc_ldpc_b200 -- B200 (sm_100a) SPARC-AMP + outer-LDPC decoder behind the entry points of Spimp/sparc_ldpc.

Layout: csrc/ (CUDA kernels + C ABI, built into libsparc_b200.so), _lib.py (ctypes binding), engine.py (device
handles), decoder.py (batched link decoders), ldpc.py / sparc_ldpc.py / amp_exit.py / amp_test.py (host-side
mirrors of the reference modules of the same name), dist.py (codeword sharding over ranks).
"""
__version__ = "0.1.0"
