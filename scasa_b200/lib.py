"""ctypes binding of libscasa_cuda.so (include/scasa_cuda.h).

No fallback: if the shared library is missing, or there is no CUDA device, every entry
point raises.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libscasa_cuda.so")
_lib = None


class ScasaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libscasa_cuda error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("maxiter_x", C.c_int32), ("maxiter_bt", C.c_int32), ("maxerr", C.c_double),
                ("lim", C.c_double), ("modify", C.c_int32), ("adjust_esp", C.c_double),
                ("hamming_dist", C.c_int32), ("fixed_iter", C.c_int32), ("fp32", C.c_int32),
                ("maxerr_bt", C.c_double), ("lim_bt", C.c_double)]


class XMat(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("q_off", C.POINTER(C.c_int32)), ("p_off", C.POINTER(C.c_int32)),
                ("x_off", C.POINTER(C.c_int64)), ("x", C.POINTER(C.c_double))]


class Crp(C.Structure):
    _fields_ = [("m", XMat), ("col_tx", C.POINTER(C.c_int32)), ("row_pat", C.POINTER(C.c_uint8)),
                ("name_rank", C.POINTER(C.c_int32))]


class FilesTiming(C.Structure):
    _fields_ = [("quant_s", C.c_double), ("transpose_s", C.c_double), ("write_iso_s", C.c_double), ("write_gene_s", C.c_double),
                ("iso_bytes", C.c_int64), ("gene_bytes", C.c_int64), ("iso_rows", C.c_int64), ("gene_rows", C.c_int64)]


class Csr(C.Structure):
    """scasa_csr_t: library-allocated sparse rows (scasa_quant_csr); release with scasa_csr_free."""
    _fields_ = [("n_rows", C.c_int64), ("nnz", C.c_int64), ("ptr", C.POINTER(C.c_int64)), ("col", C.POINTER(C.c_int32)),
                ("val", C.POINTER(C.c_double))]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("pack_ms", C.c_float), ("tasks_ms", C.c_float), ("em_ms", C.c_float),
                ("finish_ms", C.c_float), ("gene_ms", C.c_float), ("em_class_ms", C.c_float * 4),
                ("em_class_tasks", C.c_int64 * 4), ("n_tasks", C.c_int64), ("n_entries", C.c_int64),
                ("n_active", C.c_int64), ("inner_iters", C.c_int64), ("outer_iters", C.c_int64),
                ("alg_bytes", C.c_double), ("em_alg_bytes", C.c_double), ("launches", C.c_int32),
                ("max_slot_bytes", C.c_int32), ("run_parts", C.c_int32), ("em_pair_ms", C.c_float),
                ("em_pair_tasks", C.c_int64), ("result_nnz", C.c_int64)]

    def as_dict(self) -> dict:
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


# every symbol include/scasa_cuda.h declares (tests check the library exports all of them)
SYMBOLS = (
    "scasa_opts_default", "scasa_ctx_create", "scasa_ctx_destroy", "scasa_last_error", "scasa_set_opts",
    "scasa_n_rows", "scasa_n_cols", "scasa_n_em_blocks", "scasa_em_blocks", "scasa_poor_blocks",
    "scasa_chunk_split", "scasa_estimate", "scasa_stage_counts", "scasa_stage_eqcounts", "scasa_run",
    "scasa_fetch_iso", "scasa_fetch_iters", "scasa_fetch_colsum", "scasa_fetch_counts", "scasa_set_gene_map",
    "scasa_fetch_gene", "scasa_gene_sum", "scasa_eqclass_map", "scasa_last_timing", "scasa_host_alloc",
    "scasa_host_free", "scasa_device_count", "scasa_version", "scasa_set_host_threads", "scasa_expand_pairs", "scasa_quant", "scasa_transfer_bytes", "scasa_bfh_open", "scasa_bfh_close", "scasa_bfh_sizes", "scasa_bfh_read", "scasa_bus_open", "scasa_write_table", "scasa_format_real15", "scasa_set_run_parts", "scasa_eqclass_map_host",
    "scasa_result_nnz", "scasa_fetch_iso_csr", "scasa_fetch_gene_csr", "scasa_quant_multi", "scasa_write_table_sparse", "scasa_csr_transpose", "scasa_quant_files", "scasa_dm0_is_fixed_point",
    "scasa_quant_csr", "scasa_csr_free",
)


def build(force: bool = False) -> str:
    """Compile the library in-tree (nvcc, sm_100a).  Needs no GPU."""
    src = os.path.join(_HERE, "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    deps.append(os.path.join(_HERE, "..", "include", "scasa_cuda.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(SO_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["sh", os.path.join(src, "build.sh")])
    return SO_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ScasaError(-2, f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, i32p, i64p, f64p, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.scasa_last_error.restype = C.c_char_p
    L.scasa_last_error.argtypes = [vp]
    L.scasa_version.restype = C.c_char_p
    L.scasa_ctx_create.argtypes = [C.POINTER(XMat), C.POINTER(Crp), C.POINTER(Opts), C.c_int, C.POINTER(vp)]
    L.scasa_ctx_destroy.argtypes = [vp]
    L.scasa_ctx_destroy.restype = None
    L.scasa_set_opts.argtypes = [vp, C.POINTER(Opts)]
    for f in ("scasa_n_rows", "scasa_n_cols"):
        getattr(L, f).restype = C.c_int64
        getattr(L, f).argtypes = [vp]
    L.scasa_n_em_blocks.argtypes = [vp]
    L.scasa_em_blocks.argtypes = [vp, i32p]
    L.scasa_poor_blocks.argtypes = [vp, u8p]
    L.scasa_chunk_split.argtypes = [C.c_int64, C.c_int32, i64p, C.c_int64, i32p]
    L.scasa_estimate.argtypes = [vp, C.c_int64, C.c_int32, i64p, i64p, i32p, f64p, f64p, i32p]
    L.scasa_stage_counts.argtypes = [vp, C.c_int64, C.c_int32, i64p, i64p, i32p, f64p]
    L.scasa_stage_eqcounts.argtypes = [vp, C.c_int64, C.c_int32, i64p, i64p, i32p, i32p, C.c_int64, i32p]
    L.scasa_run.argtypes = [vp]
    L.scasa_set_run_parts.argtypes = [vp, C.c_int32]
    L.scasa_write_table.argtypes = [C.c_char_p, f64p, C.c_int64, C.c_int64, u8p, C.c_char_p, C.c_char_p, C.c_int32]
    L.scasa_write_table_sparse.argtypes = [C.c_char_p, C.c_int64, C.c_int64, i64p, i32p, f64p, C.c_int64, i32p, C.c_char_p, C.c_char_p,
                                           C.c_int32, i64p]
    L.scasa_csr_transpose.argtypes = [C.c_int64, C.c_int64, i64p, i32p, f64p, i64p, i32p, f64p, C.c_int32]
    L.scasa_format_real15.argtypes = [C.c_double, C.c_char_p, C.c_int32]
    L.scasa_bfh_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(vp)]
    L.scasa_bfh_close.argtypes = [vp]
    L.scasa_bus_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(vp)]
    L.scasa_bfh_close.restype = None
    L.scasa_bfh_sizes.argtypes = [vp] + [i64p] * 7
    L.scasa_bfh_read.argtypes = [vp, C.c_char_p, C.c_char_p, i64p, i32p, i64p, i32p, i32p]
    L.scasa_transfer_bytes.argtypes = [vp, i64p, i64p, C.c_int32]
    L.scasa_quant.argtypes = [vp, C.c_int64, C.c_int32, i64p, i64p, i32p, i32p, C.c_int64, i32p, f64p, f64p, f64p, i32p, C.c_int32]
    L.scasa_quant_multi.argtypes = [C.POINTER(vp), C.c_int32, C.c_int64, C.c_int32, i64p, i64p, i32p, i32p, C.c_int64, i32p, f64p, f64p, f64p,
                                    i32p, C.c_int32]
    L.scasa_quant_files.argtypes = [C.POINTER(vp), C.c_int32, C.c_int64, C.c_int32, i64p, i64p, i32p, i32p, C.c_int64, i32p, C.c_char_p,
                                    C.c_char_p, C.c_char_p, C.c_char_p, i32p, C.c_char_p, C.c_int32, f64p, C.POINTER(FilesTiming)]
    L.scasa_quant_csr.argtypes = [C.POINTER(vp), C.c_int32, C.c_int64, C.c_int32, i64p, i64p, i32p, i32p, C.c_int64, i32p, C.c_int32,
                                  C.POINTER(Csr), C.POINTER(Csr), f64p, i32p]
    L.scasa_csr_free.argtypes = [C.POINTER(Csr)]
    L.scasa_csr_free.restype = None
    L.scasa_fetch_iso.argtypes = [vp, f64p]
    L.scasa_result_nnz.argtypes = [vp, i64p, i64p]
    L.scasa_fetch_iso_csr.argtypes = [vp, i64p, i32p, f64p]
    L.scasa_fetch_gene_csr.argtypes = [vp, i64p, i32p, f64p]
    L.scasa_fetch_iters.argtypes = [vp, i32p]
    L.scasa_fetch_colsum.argtypes = [vp, f64p]
    L.scasa_fetch_counts.argtypes = [vp, i64p, i32p, f64p, C.c_int64, i64p]
    L.scasa_set_gene_map.argtypes = [vp, i32p, C.c_int32]
    L.scasa_fetch_gene.argtypes = [vp, f64p]
    L.scasa_gene_sum.argtypes = [vp, i32p, C.c_int32, f64p, C.c_int64, f64p]
    L.scasa_eqclass_map.argtypes = [vp, C.c_int64, i64p, i32p, i32p, C.c_int32]
    L.scasa_eqclass_map_host.argtypes = [C.POINTER(Crp), C.c_int32, C.c_int64, i64p, i32p, i32p, C.c_int32]
    L.scasa_dm0_is_fixed_point.argtypes = [C.POINTER(XMat), C.POINTER(XMat), i32p, C.c_double, C.c_int32, i32p, i64p]
    L.scasa_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.scasa_set_host_threads.argtypes = [vp, C.c_int32]
    L.scasa_expand_pairs.argtypes = [f64p, C.c_int64, C.c_int64, i64p, i32p, i32p, f64p, C.c_int32]
    L.scasa_host_alloc.restype = vp
    L.scasa_host_alloc.argtypes = [C.c_int64]
    L.scasa_host_free.argtypes = [vp]
    L.scasa_host_free.restype = None
    _lib = L
    return L


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_opts(**kw) -> Opts:
    o = Opts()
    load().scasa_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over cudaHostAlloc'ed (pinned) memory."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(n * dtype.itemsize, 1)
    p = load().scasa_host_alloc(nbytes)
    if not p:
        raise ScasaError(-3, f"cudaHostAlloc({nbytes}) failed")
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED: dict = {}


def pinned_free(arr: np.ndarray) -> None:
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        load().scasa_host_free(p)
