"""ctypes binding of ``libwsage.so`` (the C ABI in ``include/wsage.h``).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, the
caller gets a ``RuntimeError``.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C scdeepsort_b200/csrc``.
"""
import ctypes
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libwsage.so"

OK, EINVAL, EUNSUPPORTED, ECUDA = 0, 1, 2, 3
COL_I32, COL_U16 = 32, 16
ALGO_AUTO, ALGO_GATHER, ALGO_TILED = 0, 1, 2
D16_F16X2, D16_BF16 = 0, 1
SPLIT_ROWS, SPLIT_TRANSPOSED, SPLIT_COLBLOCKS, SPLIT_KBLOCKS, SPLIT_BLOCKED = 0, 1, 2, 3, 4

EXPORTS = ("wsage_version", "wsage_last_error", "wsage_launch_count", "wsage_block_agg_fwd",
           "wsage_block_agg_bwd", "wsage_spmm_workspace_bytes", "wsage_spmm_algo", "wsage_spmm", "wsage_amax", "wsage_split16",
           "wsage_split16_masked", "wsage_split16_colsum", "wsage_sum_slabs", "wsage_colsum_masked", "wsage_rowdot",
           "wsage_dense16_slots_pad", "wsage_dense16_splits", "wsage_dense16",
           "wsage_split_tf32", "wsage_linear_tc", "wsage_grad_w_splits", "wsage_grad_w_tc", "wsage_sample_neighbors",
           "wsage_softmax_ce", "wsage_adam_step",
           "wsage_peer_bytes", "wsage_peer_alloc", "wsage_peer_open", "wsage_peer_close", "wsage_peer_free", "wsage_peer_status",
           "wsage_peer_reduce")
PEER_MAX = 8


class SpmmArgs(Structure):
    """Mirror of ``wsage_spmm_args`` (include/wsage.h)."""
    _fields_ = [
        ("rowptr", c_void_p), ("col", c_void_p), ("col_bits", c_int32), ("x", c_void_p), ("nnz", c_int64),
        ("hs", c_void_p), ("ld_hs", c_int64), ("n_src", c_int64), ("n_dst", c_int64), ("dim", c_int32),
        ("dscale", c_void_p), ("selfcoef", c_void_p), ("hself", c_void_p), ("ld_hself", c_int64),
        ("out", c_void_p), ("ld_out", c_int64), ("raw", c_void_p), ("ld_raw", c_int64),
        ("q", c_void_p), ("ld_q", c_int64), ("dot", c_void_p), ("row_perm", c_void_p),
        ("algo", c_int32), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("init", c_void_p), ("init_slabs", c_int32), ("init_rows", c_int64), ("init_map", c_void_p),
    ]


class Dense16Args(Structure):
    """Mirror of ``wsage_dense16_args`` (include/wsage.h)."""
    _fields_ = [
        ("x_hi", c_void_p), ("x_lo", c_void_p), ("fmt", c_int32), ("cells", c_int64), ("gene_slots", c_int32),
        ("x_scale", c_float), ("side", c_int32), ("h_hi", c_void_p), ("h_lo", c_void_p), ("ld_h", c_int64),
        ("h_amax", c_void_p), ("dim", c_int32), ("n_dst", c_int64), ("n_src_cells", c_int64),
        ("dscale", c_void_p), ("selfcoef", c_void_p), ("hself", c_void_p), ("ld_hself", c_int64),
        ("out", c_void_p), ("ld_out", c_int64), ("chunk_rows", c_int32),
        ("x_amax", c_void_p), ("bias", c_void_p), ("relu", c_int32), ("deterministic", c_int32),
    ]


class PeerReduceArgs(Structure):
    """Mirror of ``wsage_peer_reduce_args`` (include/wsage.h)."""
    _fields_ = [
        ("rank", c_int32), ("world", c_int32), ("bases", POINTER(c_void_p)), ("max_elems", c_int64), ("epoch", ctypes.c_uint32),
        ("slabs", c_void_p), ("n_slabs", c_int32), ("slab_rows", c_int64), ("slot_of_row", c_void_p), ("rows", c_int64),
        ("dim", c_int32), ("dscale", c_void_p), ("selfcoef", c_void_p), ("hself", c_void_p), ("ld_hself", c_int64),
        ("out", c_void_p), ("ld_out", c_int64), ("raw", c_void_p), ("ld_raw", c_int64), ("timeout_s", c_float), ("grid", c_int32),
    ]


_lib = None


def build(verbose=False):
    """Compile libwsage.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", str(CSRC)], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"building libwsage.so failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stdout)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `make -C scdeepsort_b200/csrc` or __graft_entry__.build()). "
            "scdeepsort_b200 has no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.wsage_version.restype = c_int32
    lib.wsage_last_error.restype = c_char_p
    lib.wsage_launch_count.restype = c_int64
    lib.wsage_launch_count.argtypes = [c_int32]
    lib.wsage_block_agg_fwd.restype = c_int32
    lib.wsage_block_agg_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]
    lib.wsage_block_agg_bwd.restype = c_int32
    lib.wsage_block_agg_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                        c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                        c_void_p, c_int64, c_void_p, c_void_p]
    lib.wsage_spmm_workspace_bytes.restype = c_size_t
    lib.wsage_spmm_workspace_bytes.argtypes = [POINTER(SpmmArgs)]
    lib.wsage_spmm_algo.restype = c_int32
    lib.wsage_spmm_algo.argtypes = [POINTER(SpmmArgs)]
    lib.wsage_spmm.restype = c_int32
    lib.wsage_spmm.argtypes = [POINTER(SpmmArgs), c_void_p]
    lib.wsage_amax.restype = c_int32
    lib.wsage_amax.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.wsage_split16.restype = c_int32
    lib.wsage_split16.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32,
                                  c_void_p, c_void_p, c_int64, c_void_p]
    lib.wsage_split16_masked.restype = c_int32
    lib.wsage_split16_masked.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32,
                                         c_int32, c_void_p, c_void_p, c_int64, c_void_p]
    lib.wsage_split16_colsum.restype = c_int32
    lib.wsage_split16_colsum.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                         c_int64, c_void_p, c_int32, c_void_p, c_void_p]
    lib.wsage_sum_slabs.restype = c_int32
    lib.wsage_sum_slabs.argtypes = [c_void_p, c_int32, c_int64, c_int64, c_int32, c_void_p, c_int64, c_void_p]
    lib.wsage_colsum_masked.restype = c_int32
    lib.wsage_colsum_masked.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p]
    lib.wsage_rowdot.restype = c_int32
    lib.wsage_rowdot.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p]
    lib.wsage_dense16_slots_pad.restype = c_int32
    lib.wsage_dense16_slots_pad.argtypes = [c_int32]
    lib.wsage_dense16_splits.restype = c_int32
    lib.wsage_dense16_splits.argtypes = [POINTER(Dense16Args)]
    lib.wsage_dense16.restype = c_int32
    lib.wsage_dense16.argtypes = [POINTER(Dense16Args), c_void_p]
    lib.wsage_split_tf32.restype = c_int32
    lib.wsage_split_tf32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                     c_void_p, c_int64, c_int64, c_int32, c_void_p]
    lib.wsage_linear_tc.restype = c_int32
    lib.wsage_linear_tc.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int32,
                                    c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p]
    lib.wsage_grad_w_splits.restype = c_int32
    lib.wsage_grad_w_splits.argtypes = [c_int64, c_int32]
    lib.wsage_grad_w_tc.restype = c_int32
    lib.wsage_grad_w_tc.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                    c_void_p, c_int32, c_void_p, c_int64, c_void_p]
    lib.wsage_sample_neighbors.restype = c_int32
    lib.wsage_sample_neighbors.argtypes = [c_void_p, c_void_p, c_int64, c_int32, ctypes.c_uint64, c_void_p, c_void_p, c_void_p]
    lib.wsage_softmax_ce.restype = c_int32
    lib.wsage_softmax_ce.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_int32, c_void_p]
    lib.wsage_adam_step.restype = c_int32
    lib.wsage_adam_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_double, c_int32, c_void_p]
    lib.wsage_peer_bytes.restype = c_size_t
    lib.wsage_peer_bytes.argtypes = [c_int64]
    lib.wsage_peer_alloc.restype = c_int32
    lib.wsage_peer_alloc.argtypes = [c_int64, POINTER(c_void_p), c_void_p]
    lib.wsage_peer_open.restype = c_int32
    lib.wsage_peer_open.argtypes = [c_void_p, POINTER(c_void_p)]
    lib.wsage_peer_close.restype = c_int32
    lib.wsage_peer_close.argtypes = [c_void_p]
    lib.wsage_peer_free.restype = c_int32
    lib.wsage_peer_free.argtypes = [c_void_p]
    lib.wsage_peer_status.restype = c_int32
    lib.wsage_peer_status.argtypes = [c_void_p, POINTER(c_int32)]
    lib.wsage_peer_reduce.restype = c_int32
    lib.wsage_peer_reduce.argtypes = [POINTER(PeerReduceArgs), c_void_p]
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        msg = load().wsage_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")


def launch_count(reset=False):
    return int(load().wsage_launch_count(1 if reset else 0))
