"""ctypes binding of libagdiff_b200.so (C ABI in include/agdiff_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, the caller
gets an exception.  ``load()`` only dlopens the library (possible without a GPU, used by the CPU
test-suite to check the exported symbols); every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libagdiff_b200.so")

AGD_OK, AGD_ERR_INVALID, AGD_ERR_CUDA, AGD_ERR_CAPACITY, AGD_ERR_NAN, AGD_ERR_RANGE = 0, -1, -2, -3, -4, -5
MODE_FFMA, MODE_TF32, MODE_F16 = 0, 1, 2
MAX_MOL_ATOMS = 512
MAX_RADIUS_NBRS = 32

EXPORTS = [
    "agd_abi_version", "agd_last_error", "agd_create", "agd_destroy", "agd_weight_slot_count",
    "agd_weight_slot_name", "agd_weight_slot_size", "agd_load_weights", "agd_batch_workspace_bytes",
    "agd_batch_create", "agd_batch_destroy", "agd_build_edges", "agd_forward", "agd_sample",
    "agd_extend_bond_order", "agd_op_cfconv_aggregate", "agd_op_eq_transform", "agd_op_gin_message", "agd_op_eq_transform_segments", "agd_op_kabsch_rmsd", "agd_host_local_pairs",
    "agd_debug_fetch",
    "agd_launch_count", "agd_profile_forward", "agd_forward_edges",
    "agd_nan_steps", "agd_set_mode", "agd_get_mode", "agd_set_option", "agd_range_flag", "agd_f16_lo_shift", "agd_debug_timing",
]


class AgdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("agdiff_b200 native error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("hidden_dim", C.c_int32), ("num_convs", C.c_int32), ("num_convs_local", C.c_int32),
                ("smooth_conv", C.c_int32), ("cutoff", C.c_float), ("device", C.c_int32)]


class BatchDesc(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("n_mols", C.c_int32),
                ("atom_type", C.c_void_p), ("mol_ptr", C.c_void_p), ("atom_mol", C.c_void_p), ("mol_gid", C.c_void_p),
                ("n_static", C.c_int32),
                ("st_src", C.c_void_p), ("st_dst", C.c_void_p), ("st_type", C.c_void_p), ("st_in_ptr", C.c_void_p),
                ("n_local", C.c_int32),
                ("lc_src", C.c_void_p), ("lc_dst", C.c_void_p), ("lc_type", C.c_void_p), ("lc_in_ptr", C.c_void_p),
                ("lc_canon", C.c_void_p), ("lc_out_ptr", C.c_void_p), ("lc_cdst", C.c_void_p),
                ("edge_capacity", C.c_int64)]


class SampleParams(C.Structure):
    _fields_ = [("n_steps", C.c_int32),
                ("sigma", C.c_void_p), ("step_size", C.c_void_p), ("noise_scale", C.c_void_p), ("use_global", C.c_void_p),
                ("w_global", C.c_float), ("clip", C.c_float), ("clip_local", C.c_float), ("clip_pos", C.c_float),
                ("seed", C.c_uint64), ("noise", C.c_void_p), ("traj", C.c_void_p), ("use_cuda_graph", C.c_int32),
                ("step_offset", C.c_int32)]


class EdgeSet(C.Structure):
    _fields_ = [("n_edges", C.c_int32),
                ("e_src", C.c_void_p), ("e_dst", C.c_void_p), ("e_type", C.c_void_p), ("e_canon", C.c_void_p),
                ("e_len", C.c_void_p), ("in_ptr", C.c_void_p), ("out_ptr", C.c_void_p),
                ("c_src", C.c_void_p), ("c_dst", C.c_void_p), ("c_type", C.c_void_p), ("c_len", C.c_void_p),
                ("lc_len", C.c_void_p)]


class ForwardOut(C.Structure):
    _fields_ = [("edge_inv_global", C.c_void_p), ("edge_inv_local", C.c_void_p), ("edge_row", C.c_void_p),
                ("edge_col", C.c_void_p), ("edge_type", C.c_void_p), ("edge_length", C.c_void_p), ("n_edges", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m agdiff_b200.build` "
                          "(agdiff_b200 has no CPU or PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.agd_abi_version.restype = C.c_int
    lib.agd_last_error.restype = C.c_char_p
    lib.agd_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.agd_destroy.argtypes = [vp]
    lib.agd_destroy.restype = None
    lib.agd_weight_slot_count.argtypes = [vp]
    lib.agd_weight_slot_name.argtypes = [vp, C.c_int]
    lib.agd_weight_slot_name.restype = C.c_char_p
    lib.agd_weight_slot_size.argtypes = [vp, C.c_int]
    lib.agd_weight_slot_size.restype = i64
    lib.agd_load_weights.argtypes = [vp, vp, vp, C.c_int, i64]
    lib.agd_batch_workspace_bytes.argtypes = [vp, C.POINTER(BatchDesc)]
    lib.agd_batch_workspace_bytes.restype = i64
    lib.agd_batch_create.argtypes = [vp, C.POINTER(BatchDesc), C.POINTER(vp)]
    lib.agd_batch_destroy.argtypes = [vp]
    lib.agd_batch_destroy.restype = None
    lib.agd_build_edges.argtypes = [vp, vp, vp, C.POINTER(ForwardOut), vp]
    lib.agd_forward.argtypes = [vp, vp, vp, C.POINTER(ForwardOut), vp]
    lib.agd_forward_edges.argtypes = [vp, vp, vp, C.POINTER(EdgeSet), C.POINTER(ForwardOut), vp]
    lib.agd_sample.argtypes = [vp, vp, vp, C.POINTER(SampleParams), C.POINTER(i32), vp]
    lib.agd_extend_bond_order.argtypes = [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.agd_op_cfconv_aggregate.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.agd_op_eq_transform.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp, vp]
    lib.agd_op_gin_message.argtypes = [vp, vp, vp, vp, i32, C.c_float, vp, vp]
    lib.agd_op_eq_transform_segments.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.agd_op_kabsch_rmsd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.agd_host_local_pairs.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.agd_debug_fetch.argtypes = [vp, C.c_char_p, vp, i64]
    lib.agd_debug_fetch.restype = i64
    lib.agd_profile_forward.argtypes = [vp, vp, vp, i32, C.c_char_p, i64, vp, i32, C.POINTER(i32), C.POINTER(i32), vp]
    lib.agd_launch_count.argtypes = [vp]
    lib.agd_launch_count.restype = i64
    lib.agd_set_mode.argtypes = [vp, C.c_int]
    lib.agd_get_mode.argtypes = [vp]
    lib.agd_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.agd_nan_steps.argtypes = [vp, vp, C.c_int32]
    lib.agd_nan_steps.restype = C.c_int
    lib.agd_range_flag.argtypes = [vp, C.POINTER(i32)]
    lib.agd_f16_lo_shift.restype = C.c_int
    lib.agd_debug_timing.argtypes = [vp, vp]
    if lib.agd_abi_version() != 1:
        raise ImportError("libagdiff_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != AGD_OK:
        msg = load().agd_last_error()
        raise AgdError(code, msg.decode() if msg else "")
