# SPDX-License-Identifier: Apache-2.0
"""ctypes binding of libxntt.so (include/xntt.h) - plumbing for tests/ and bench.py.

The product is the C-ABI library plus the C++20 front-end in ``host/sventt``; this module only
hands raw device pointers (e.g. ``torch.Tensor.data_ptr()``) and stream handles to it.  There is no
CPU path here: if the CUDA library has not been built, importing succeeds but ``load()`` raises.

The directory name carries a hyphen, so load it with ``__graft_entry__.load_package()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxntt.so")

P0 = 0xFFFFFC6E80000001  # 2^64 - 1827*2^31 + 1 (reference README.md:19)
G0 = 3

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_ALLOC, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4, -5
ENABLE_FORWARD, ENABLE_INVERSE, COMPACT_TABLES, MODMUL_FIXED_POINT = 1, 2, 4, 8
TILES_WIDE, TILES_NARROW = 16, 32
MAX_SPLITS = 4


class XnttError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: status {status}" + (f" ({detail})" if detail else ""))


class Desc(C.Structure):
    _fields_ = [
        ("modulus", C.c_uint64),
        ("generator", C.c_uint64),
        ("log2_m", C.c_uint32),
        ("batch", C.c_uint32),
        ("inverse_factor", C.c_uint64),
        ("flags", C.c_uint32),
        ("device", C.c_int32),
        ("n_splits", C.c_uint32),
        ("split_log2", C.c_uint32 * MAX_SPLITS),
        ("shard_count", C.c_uint32),
        ("shard_rank", C.c_uint32),
        ("twist_table_max_mb", C.c_uint32),
        ("reserved_", C.c_uint32),
    ]


# every symbol include/xntt.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U64P = C.c_void_p  # raw addresses (device or host)
SYMBOLS = {
    "xntt_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(Desc)]),
    "xntt_plan_destroy": (C.c_int, [_P]),
    "xntt_plan_m": (C.c_uint64, [_P]),
    "xntt_plan_batch": (C.c_uint32, [_P]),
    "xntt_plan_launches": (C.c_uint32, [_P, C.c_int]),
    "xntt_plan_modmul": (C.c_uint32, [_P]),
    "xntt_plan_twiddle_form": (C.c_uint32, [_P, C.c_uint32, C.c_int]),
    "xntt_plan_tile_log2": (C.c_uint32, [_P, C.c_uint32]),
    "xntt_plan_splits": (C.c_uint32, [_P, C.POINTER(C.c_uint32), C.c_uint32]),
    "xntt_forward": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_inverse": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_forward_multiply": (C.c_int, [_P, _U64P, _U64P, _U64P, _P]),
    "xntt_run_pass": (C.c_int, [_P, C.c_uint32, C.c_int, _U64P, _U64P, _P]),
    "xntt_forward_host": (C.c_int, [_P, _U64P, _U64P]),
    "xntt_inverse_host": (C.c_int, [_P, _U64P, _U64P]),
    "xntt_shard_forward_cols": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_shard_forward_rows": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_shard_inverse_rows": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_shard_inverse_cols": (C.c_int, [_P, _U64P, _U64P, _P]),
    "xntt_shard_forward_cols_chunk": (C.c_int, [_P, _U64P, _U64P, C.c_uint32, C.c_uint32, _P]),
    "xntt_shard_forward_rows_tiled": (C.c_int, [_P, _U64P, _U64P, C.c_uint32, _P]),
    "xntt_shard_inverse_rows_tiled": (C.c_int, [_P, _U64P, _U64P, _U64P, C.c_uint32, _P]),
    "xntt_shard_inverse_cols_chunk": (C.c_int, [_P, _U64P, _U64P, C.c_uint32, C.c_uint32, _P]),
    "xntt_shard_forward_cols_peer": (C.c_int, [_P, C.POINTER(C.c_void_p), _U64P, _P]),
    "xntt_shard_inverse_rows_peer": (C.c_int, [_P, C.POINTER(C.c_void_p), _U64P, _U64P, _P]),
    "xntt_mgpu_create": (C.c_int, [C.POINTER(_P), C.POINTER(Desc), C.POINTER(C.c_int32), C.c_uint32]),
    "xntt_mgpu_destroy": (C.c_int, [_P]),
    "xntt_mgpu_devices": (C.c_uint32, [_P]),
    "xntt_mgpu_m": (C.c_uint64, [_P]),
    "xntt_mgpu_n0": (C.c_uint64, [_P]),
    "xntt_mgpu_stream": (C.c_void_p, [_P, C.c_uint32]),
    "xntt_mgpu_forward": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "xntt_mgpu_inverse": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "xntt_mgpu_synchronize": (C.c_int, [_P]),
    "xntt_mgpu_forward_host": (C.c_int, [_P, _U64P, _U64P]),
    "xntt_mgpu_inverse_host": (C.c_int, [_P, _U64P, _U64P]),
    "xntt_to_montgomery": (C.c_int, [_P, _U64P, _U64P, C.c_size_t, _P]),
    "xntt_from_montgomery": (C.c_int, [_P, _U64P, _U64P, C.c_size_t, _P]),
    "xntt_multiply_normalize": (C.c_int, [_P, _U64P, _U64P, _U64P, C.c_size_t, _P]),
    "xntt_transpose": (C.c_int, [_U64P, _U64P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "xntt_alloc_device": (C.c_int, [C.POINTER(_P), C.c_size_t, C.c_int]),
    "xntt_free_device": (C.c_int, [_P]),
    "xntt_alloc_pinned": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "xntt_free_pinned": (C.c_int, [_P]),
    "xntt_memcpy_h2d": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "xntt_memcpy_d2h": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "xntt_stream_synchronize": (C.c_int, [_P]),
    "xntt_pointer_is_device": (C.c_int, [_P]),
    "xntt_strerror": (C.c_char_p, [C.c_int]),
    "xntt_last_cuda_error": (C.c_char_p, []),
    "xntt_version": (C.c_char_p, []),
    "xntt_device_count": (C.c_int, []),
    "xntt_kinnaes_sum": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                 C.POINTER(C.c_uint64)]),
    "xntt_kinnaes_compute": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "xntt_release_scratch": (C.c_int, []),
    "xntt_microbench": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


class Library:
    """A loaded libxntt (or, in the CPU test-suite only, the host emulator built from the same
    sources)."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args

    def check(self, status, what):
        if status != OK:
            detail = self.lib.xntt_strerror(status).decode()
            if status == ERR_CUDA or status == ERR_ALLOC:
                detail += ": " + self.lib.xntt_last_cuda_error().decode()
            raise XnttError(status, what, detail)

    def version(self):
        return self.lib.xntt_version().decode()

    def device_count(self):
        return self.lib.xntt_device_count()

    def plan(self, log2_m, **kw):
        return Plan(self, log2_m, **kw)

    def mgpu(self, log2_m, devices, **kw):
        return MultiGpu(self, log2_m, devices, **kw)

    def transpose(self, dst, src, rows, cols, ld_dst=None, ld_src=None, stream=0):
        self.check(self.lib.xntt_transpose(dst, src, rows, cols, rows if ld_dst is None else ld_dst,
                                           cols if ld_src is None else ld_src, stream), "xntt_transpose")

    # MagicSeriesKinnaes<m, PAdic64<Modulus<modulus, generator>>, n> (examples/magic-series-kinnaes/kinnaes.hpp)
    def kinnaes_sum(self, modulus, generator, m, n, j_begin, j_end, device=-1):
        r = C.c_uint64()
        self.check(self.lib.xntt_kinnaes_sum(modulus, generator, m, n, j_begin, j_end, device, C.byref(r)),
                   "xntt_kinnaes_sum")
        return int(r.value)

    def kinnaes_compute(self, modulus, generator, m, n, device=-1):
        r = C.c_uint64()
        self.check(self.lib.xntt_kinnaes_compute(modulus, generator, m, n, device, C.byref(r)), "xntt_kinnaes_compute")
        return int(r.value)

    def microbench(self, kind, iters):
        g, ms = C.c_double(), C.c_double()
        self.check(self.lib.xntt_microbench(kind, iters, C.byref(g), C.byref(ms)), "xntt_microbench")
        return g.value, ms.value


class Plan:
    """sventt::NTT<kernel> (include/sventt/wrapper.hpp:13-83) over raw pointers."""

    def __init__(self, library, log2_m, modulus=P0, generator=G0, batch=1, inverse_factor=None,
                 forward=True, inverse=True, device=-1, splits=None, shard_count=0, shard_rank=0,
                 compact_tables=False, twist_table_max_mb=0, fixed_point=False, tiles=None):
        """tiles: None = the planner's size rule, "wide" / "narrow" = XNTT_TILES_WIDE / XNTT_TILES_NARROW"""
        self.L = library
        d = Desc()
        d.modulus, d.generator = modulus, generator
        d.log2_m, d.batch = log2_m, batch
        d.inverse_factor = (1 << log2_m) if inverse_factor is None else inverse_factor
        d.flags = (ENABLE_FORWARD if forward else 0) | (ENABLE_INVERSE if inverse else 0) | \
            (COMPACT_TABLES if compact_tables else 0) | (MODMUL_FIXED_POINT if fixed_point else 0) | \
            {None: 0, "wide": TILES_WIDE, "narrow": TILES_NARROW}[tiles]
        d.device = device
        if splits:
            d.n_splits = len(splits)
            for i, s in enumerate(splits):
                d.split_log2[i] = s
        d.shard_count, d.shard_rank = shard_count, shard_rank
        d.twist_table_max_mb = twist_table_max_mb
        self.desc = d
        self.h = C.c_void_p()
        library.check(library.lib.xntt_plan_create(C.byref(self.h), C.byref(d)), "xntt_plan_create")
        self.m = library.lib.xntt_plan_m(self.h)
        self.batch = library.lib.xntt_plan_batch(self.h)

    def close(self):
        if self.h:
            self.L.lib.xntt_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def splits(self):
        buf = (C.c_uint32 * MAX_SPLITS)()
        n = self.L.lib.xntt_plan_splits(self.h, buf, MAX_SPLITS)
        return [int(buf[i]) for i in range(n)]

    @property
    def tile_log2(self):
        """log2 of the residues one CTA works on, per pass"""
        return [int(self.L.lib.xntt_plan_tile_log2(self.h, i)) for i in range(len(self.splits))]

    def twiddle_forms(self, inverse=False):
        """per pass: 0 row pass, 1 compact tables, 2 whole matrix by the pass itself, 3 whole matrix by its neighbour"""
        return [int(self.L.lib.xntt_plan_twiddle_form(self.h, i, 1 if inverse else 0)) for i in range(len(self.splits))]

    @property
    def modmul(self):
        """0 = Montgomery (PAdic64) kernels, 1 = Shoup (FixedPoint64) kernels"""
        return int(self.L.lib.xntt_plan_modmul(self.h))

    @property
    def launches(self):
        return int(self.L.lib.xntt_plan_launches(self.h, 0))

    def _call(self, name, *args):
        self.L.check(getattr(self.L.lib, name)(self.h, *args), name)

    # device pointers (ints) + stream handle (int)
    def forward(self, dst, src, stream=0):
        self._call("xntt_forward", dst, src, stream)

    def inverse(self, dst, src, stream=0):
        self._call("xntt_inverse", dst, src, stream)

    def forward_multiply(self, dst, src, b_mont, stream=0):
        self._call("xntt_forward_multiply", dst, src, b_mont, stream)

    def run_pass(self, index, inverse, dst, src, stream=0):
        self._call("xntt_run_pass", index, 1 if inverse else 0, dst, src, stream)

    def forward_host(self, dst, src):
        self._call("xntt_forward_host", dst, src)

    def inverse_host(self, dst, src):
        self._call("xntt_inverse_host", dst, src)

    def shard_forward_cols(self, dst, src, stream=0):
        self._call("xntt_shard_forward_cols", dst, src, stream)

    def shard_forward_rows(self, dst, src, stream=0):
        self._call("xntt_shard_forward_rows", dst, src, stream)

    def shard_inverse_rows(self, dst, src, stream=0):
        self._call("xntt_shard_inverse_rows", dst, src, stream)

    def shard_inverse_cols(self, dst, src, stream=0):
        self._call("xntt_shard_inverse_cols", dst, src, stream)

    def shard_forward_cols_chunk(self, tiles, src, chunk, nchunks, stream=0):
        self._call("xntt_shard_forward_cols_chunk", tiles, src, chunk, nchunks, stream)

    def shard_forward_rows_tiled(self, dst, tiles, nchunks, stream=0):
        self._call("xntt_shard_forward_rows_tiled", dst, tiles, nchunks, stream)

    def shard_inverse_rows_tiled(self, tiles, src, work, nchunks, stream=0):
        self._call("xntt_shard_inverse_rows_tiled", tiles, src, work, nchunks, stream)

    def shard_inverse_cols_chunk(self, dst, tiles, chunk, nchunks, stream=0):
        self._call("xntt_shard_inverse_cols_chunk", dst, tiles, chunk, nchunks, stream)

    def shard_forward_cols_peer(self, peers, src, stream=0):
        arr = (C.c_void_p * len(peers))(*peers)
        self._call("xntt_shard_forward_cols_peer", arr, src, stream)

    def shard_inverse_rows_peer(self, peers, src, work, stream=0):
        arr = (C.c_void_p * len(peers))(*peers)
        self._call("xntt_shard_inverse_rows_peer", arr, src, work, stream)

    def to_montgomery(self, dst, src, count, stream=0):
        self._call("xntt_to_montgomery", dst, src, count, stream)

    def from_montgomery(self, dst, src, count, stream=0):
        self._call("xntt_from_montgomery", dst, src, count, stream)

    def multiply_normalize(self, dst, a, b_mont, count, stream=0):
        self._call("xntt_multiply_normalize", dst, a, b_mont, count, stream)


class MultiGpu:
    """xntt_mgpu: one transform over several GPUs of this process (C++-hosted exchange, include/xntt.h)."""

    def __init__(self, library, log2_m, devices, modulus=P0, generator=G0, inverse_factor=None, splits=None,
                 forward=True, inverse=True, fixed_point=False):
        self.L = library
        d = Desc()
        d.modulus, d.generator = modulus, generator
        d.log2_m, d.batch = log2_m, 1
        d.inverse_factor = (1 << log2_m) if inverse_factor is None else inverse_factor
        d.flags = (ENABLE_FORWARD if forward else 0) | (ENABLE_INVERSE if inverse else 0) | \
            (MODMUL_FIXED_POINT if fixed_point else 0)
        d.device = -1
        if splits:
            d.n_splits = len(splits)
            for i, s in enumerate(splits):
                d.split_log2[i] = s
        self.h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*devices)
        library.check(library.lib.xntt_mgpu_create(C.byref(self.h), C.byref(d), devs, len(devices)), "xntt_mgpu_create")
        self.G = len(devices)
        self.m = library.lib.xntt_mgpu_m(self.h)
        self.n0 = library.lib.xntt_mgpu_n0(self.h)
        self.n1 = self.m // self.n0

    def close(self):
        if self.h:
            self.L.lib.xntt_mgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptrs(self, ptrs):
        assert len(ptrs) == self.G
        return (C.c_void_p * self.G)(*ptrs)

    def forward(self, dst_ptrs, src_ptrs):
        self.L.check(self.L.lib.xntt_mgpu_forward(self.h, self._ptrs(dst_ptrs), self._ptrs(src_ptrs)), "xntt_mgpu_forward")

    def inverse(self, dst_ptrs, src_ptrs):
        self.L.check(self.L.lib.xntt_mgpu_inverse(self.h, self._ptrs(dst_ptrs), self._ptrs(src_ptrs)), "xntt_mgpu_inverse")

    def synchronize(self):
        self.L.check(self.L.lib.xntt_mgpu_synchronize(self.h), "xntt_mgpu_synchronize")

    def stream(self, rank):
        return self.L.lib.xntt_mgpu_stream(self.h, rank) or 0

    def forward_host(self, dst, src):
        self.L.check(self.L.lib.xntt_mgpu_forward_host(self.h, dst, src), "xntt_mgpu_forward_host")

    def inverse_host(self, dst, src):
        self.L.check(self.L.lib.xntt_mgpu_inverse_host(self.h, dst, src), "xntt_mgpu_inverse_host")


_default = None


def load():
    """The CUDA library; raises if it has not been built (no fallback)."""
    global _default
    if _default is None:
        _default = Library(LIB_PATH)
    return _default
