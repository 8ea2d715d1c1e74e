# SPDX-License-Identifier: Apache-2.0
"""ctypes view of the CPU oracle (oracle/_build/libntt_oracle.so) and, when present, of the
reference's own oracle class (oracle/_ref/libnttref.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libntt_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libnttref.so")
REF_SCALAR_SO = os.path.join(ORACLE_DIR, "_ref", "libnttref_scalar.so")

_u64 = C.c_uint64
_ptr = C.c_void_p


def _build_oracle():
    src = os.path.join(ORACLE_DIR, "ntt_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["make", "_build/libntt_oracle.so"], cwd=ORACLE_DIR, check=True)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class Oracle:
    def __init__(self):
        _build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        for name, res, args in [
            ("oracle_mulmod", _u64, [_u64, _u64, _u64]),
            ("oracle_powmod", _u64, [_u64, _u64, _u64]),
            ("oracle_addmod", _u64, [_u64, _u64, _u64]),
            ("oracle_submod", _u64, [_u64, _u64, _u64]),
            ("oracle_root_forward", _u64, [_u64, _u64, _u64]),
            ("oracle_root_inverse", _u64, [_u64, _u64, _u64]),
            ("oracle_montgomery_inverse", _u64, [_u64]),
            ("oracle_to_montgomery", _u64, [_u64, _u64]),
            ("oracle_from_montgomery", _u64, [_u64, _u64]),
            ("oracle_precompute", _u64, [_u64, _u64]),
            ("oracle_multiply_normalize", _u64, [_u64, _u64, _u64, _u64]),
            ("oracle_ntt_forward", None, [_ptr, _ptr, _u64, _u64, _u64]),
            ("oracle_ntt_inverse", None, [_ptr, _ptr, _u64, _u64, _u64]),
            ("oracle_pointwise_mul", None, [_ptr, _ptr, _ptr, _u64, _u64]),
            ("oracle_dft_point", _u64, [_ptr, _u64, _u64, _u64, _u64]),
            ("oracle_fill_xorshift", None, [_ptr, _u64, _u64, _u64]),
            ("oracle_fnv64", _u64, [_ptr, _u64]),
            ("oracle_kinnaes_comb", _u64, [_u64, _u64, _u64]),
            ("oracle_kinnaes_sum", _u64, [_u64, _u64, _u64, _u64, _u64, _u64]),
            ("oracle_kinnaes_compute", _u64, [_u64, _u64, _u64, _u64]),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    def ntt_forward(self, a, N, g):
        out = np.empty_like(a)
        self.lib.oracle_ntt_forward(_p(out), _p(a), a.size, N, g)
        return out

    def ntt_inverse(self, a, N, g):
        out = np.empty_like(a)
        self.lib.oracle_ntt_inverse(_p(out), _p(a), a.size, N, g)
        return out

    def pointwise_mul(self, a, b, N):
        out = np.empty_like(a)
        self.lib.oracle_pointwise_mul(_p(out), _p(a), _p(b), a.size, N)
        return out

    def dft_point(self, a, N, g, pos):
        return int(self.lib.oracle_dft_point(_p(a), a.size, N, g, pos))

    def fill_xorshift(self, count, seed, N):
        out = np.empty(count, dtype=np.uint64)
        self.lib.oracle_fill_xorshift(_p(out), count, seed, N)
        return out

    def kinnaes_sum(self, N, g, m, n, j_begin, j_end):
        return int(self.lib.oracle_kinnaes_sum(N, g, m, n, j_begin, j_end))

    def kinnaes_compute(self, N, g, m, n):
        return int(self.lib.oracle_kinnaes_compute(N, g, m, n))

    def fnv64(self, a):
        return int(self.lib.oracle_fnv64(_p(a), a.size))

    def __getattr__(self, name):  # scalar helpers: o.mulmod(x, y, N) ...
        fn = getattr(self.lib, "oracle_" + name)
        return lambda *args: int(fn(*args))


class Reference:
    """NTTReference / sventt::Modulus compiled from /root/reference (oracle/_ref)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        for name, res, args in [
            ("ref_ntt_forward", None, [_ptr, _ptr, _u64, _u64, _u64]),
            ("ref_ntt_inverse", None, [_ptr, _ptr, _u64, _u64, _u64]),
            ("ref_root", _u64, [C.c_int, C.c_int, _u64]),
            ("ref_montgomery_inverse", _u64, [C.c_int]),
            ("ref_multiply", _u64, [C.c_int, _u64, _u64]),
            ("ref_power", _u64, [C.c_int, _u64, _u64]),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    def ntt_forward(self, a, N, g):
        out = np.empty_like(a)
        self.lib.ref_ntt_forward(_p(out), _p(a), a.size, N, g)
        return out

    def ntt_inverse(self, a, N, g):
        out = np.empty_like(a)
        self.lib.ref_ntt_inverse(_p(out), _p(a), a.size, N, g)
        return out


class ReferenceScalar:
    """The reference's portable scalar kernel - IterativeNTT over RadixEightScalarLayer<PAdic64Scalar>
    (layer/scalar/radix-eight.hpp), compiled from /root/reference by oracle/refscalar.cpp.  Correct only below 2^62,
    hence fixed to the 62-bit test prime; outputs are lazily reduced (compare % N).  CPU baselines B2 (1 thread) and
    B3 (OpenMP over a batch) of BASELINE.md section 3."""
    SIZES = (12, 20, 24)

    def __init__(self):
        if not os.path.exists(REF_SCALAR_SO):
            raise FileNotFoundError(REF_SCALAR_SO)
        L = self.lib = C.CDLL(REF_SCALAR_SO)
        L.refscalar_modulus.restype = _u64
        L.refscalar_generator.restype = _u64
        L.refscalar_run.restype = C.c_int
        L.refscalar_run.argtypes = [C.c_int, C.c_int, _ptr, _ptr, _u64, C.c_int]
        self.N, self.g = int(L.refscalar_modulus()), int(L.refscalar_generator())

    def run(self, log2_m, inverse, a, batch=1, threads=1, out=None):
        assert a.size == batch << log2_m
        out = np.empty_like(a) if out is None else out
        rc = self.lib.refscalar_run(log2_m, 1 if inverse else 0, _p(out), _p(a), batch, threads)
        if rc != 0:
            raise ValueError(f"refscalar_run: unsupported size 2^{log2_m}")
        return out


def have_reference():
    return os.path.exists(REF_SO)


def have_reference_scalar():
    return os.path.exists(REF_SCALAR_SO)
