"""ctypes binding of libglam_b200.so (the C ABI declared in include/glam_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglam_b200.so")

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int
F32 = C.c_float
SZ = C.c_size_t

# name -> (restype, argtypes); mirrors include/glam_b200.h one to one
SIGNATURES = {
    "glam_abi_version": (I32, []),
    "glam_last_error": (C.c_char_p, []),
    "glam_launch_count": (I64, []),
    "glam_set_math_mode": (I32, [I32]),
    "glam_get_math_mode": (I32, []),
    "glam_csr_workspace_bytes": (SZ, [I64, I64]),
    "glam_build_csr": (I32, [P, I64, I64, P, P, P, P, P, P, P, P, SZ, P]),
    "glam_graph_ptr": (I32, [P, I64, I64, P, P]),
    "glam_gather_rows": (I32, [P, P, I64, I64, P, P]),
    "glam_gemm": (I32, [P, I64, P, I64, I64, P, P, I64, P, I64, I64, I64, I64, I32, P]),
    "glam_gemm_ex": (I32, [P, I64, P, I64, I64, P, P, I64, P, I64, I64, I64, I64, I32, I32, I32, P]),
    "glam_gemm_tn_workspace_bytes": (SZ, [I64, I64, I64]),
    "glam_gemm_tn": (I32, [P, I64, P, I64, I64, I64, I64, P, I64, P, SZ, P]),
    "glam_gemm_tn_ex_workspace_bytes": (SZ, [I64, I64, I64, I32]),
    "glam_gemm_tn_ex": (I32, [P, I64, P, I64, I64, I64, I64, P, I64, I32, P, P, SZ, P]),
    "glam_colsum_workspace_bytes": (SZ, [I64, I64]),
    "glam_colsum": (I32, [P, I64, I64, I64, P, P, SZ, P]),
    "glam_edge_tile_rows": (I32, [I64]),
    "glam_edge_tile_count": (I64, [I64]),
    "glam_build_edge_tiles": (I32, [P, P, P, P, I64, I64, P, P, P]),
    "glam_triplet_edge_fwd": (I32, [P, I64, P, P, P, P, P, P, I64, I64, I32, I32, I32, F32, P, P, P]),
    "glam_triplet_bwd_workspace_bytes": (SZ, [I32, I32, I32]),
    "glam_triplet_edge_bwd_dst": (I32, [P, I64, P, P, P, P, P, P, P, P, P, I64, I64, I32, I32, I32, F32, P, P, P, P, SZ, P]),
    "glam_triplet_edge_bwd_src": (I32, [P, P, P, P, P, P, P, P, P, I64, I64, I32, I32, I32, P, I64, P]),
    "glam_triplet_prep_fwd": (I32, [P, P, P, I32, I32, I32, I32, I32, P, P, P]),
    "glam_triplet_prep_bwd": (I32, [P, P, P, P, P, P, I32, I32, I32, I32, I32, P, P, P, P]),
    "glam_gru_gates_fwd": (I32, [P, P, P, P, I64, I32, I32, F32, P, P, P]),
    "glam_gru_gates_bwd": (I32, [P, P, P, P, P, P, I64, I32, I32, F32, P, P, P, P, P]),
    "glam_gru_gates_bwd_ex": (I32, [P, P, I64, P, P, P, P, I64, I32, I32, F32, P, P, P, P, P]),
    "glam_lstm_gates_fwd": (I32, [P, P, I64, I32, P, P, P]),
    "glam_lstm_gates_bwd": (I32, [P, P, P, P, P, I64, I32, P, P, P]),
    "glam_seg_attn_pool_fwd": (I32, [P, I64, P, I64, P, P, I64, I32, P, P, I64, P, P]),
    "glam_seg_attn_pool_bwd": (I32, [P, I64, P, I64, P, P, I64, P, P, I64, I32, I32, P, I64, P, P, P]),
    "glam_gru_fused_supported": (I32, [I32]),
    "glam_gru_fused_fwd": (I32, [P, I64, P, I64, P, P, P, P, P, I64, I32, I32, F32, P, P, P, P, P]),
    "glam_pool5_fwd": (I32, [P, I64, P, I64, I32, P, P, P]),
    "glam_pool5_bwd": (I32, [P, P, P, I64, I32, P, I64, P]),
    "glam_csr_aggregate": (I32, [P, I64, P, P, P, P, P, P, I64, I32, P, I64, I32, P]),
    "glam_gcn_norm": (I32, [P, P, P, P, I64, P, P, P, P]),
    "glam_adam_step": (I32, [P, P, P, P, I64, P, P, F32, F32, F32, F32, F32, P]),
    "glam_set2set_round_fwd": (I32, [P, I64, P, I64, I32, P, P, P, P, P, P, P]),
    "glam_set2set_round_bwd": (I32, [P, I64, P, I64, I32, P, P, P, P, P, I64, I32, P, P, I32, P, P]),
    "glam_pair_dot_pool_fwd": (I32, [P, P, P, P, I64, I32, P, P, P, P, P]),
    "glam_pair_dot_pool_fwd_idx": (I32, [P, P, P, P, P, I64, I32, P, P, P, P, P]),
    "glam_pair_dot_pool_fwd_tc": (I32, [P, P, P, P, P, I64, I32, P, P, P, P, P]),
    "glam_pair_dot_pool_fwd_small": (I32, [P, P, P, P, P, I64, I32, P, P, P, P, P]),
    "glam_pair_dot_pool_small_supported": (I32, [I32]),
    "glam_pair_dot_pool_tc_supported": (I32, [I32]),
    "glam_pair_dot_pool_bwd": (I32, [P, P, P, P, P, P, P, P, I64, I32, P, P, P]),
    "glam_graph_tile_caps": (I32, [P, P]),
    "glam_tile_order": (I32, [P, P, I64, I32, I32, P]),
    "glam_graph_tiles_workspace_bytes": (SZ, [I64]),
    "glam_build_graph_tiles": (I32, [P, I64, P, P, I64, I64, P, P, P, SZ, P]),
    "glam_edge_types": (I32, [P, P, I64, I32, P, P, P]),
    "glam_pair_norm_fwd": (I32, [P, I64, P, I64, I32, F32, P, I64, P]),
    "glam_pair_norm_bwd": (I32, [P, I64, P, I64, P, I64, I32, F32, P, I64, I32, P]),
    "glam_unpack_graphs": (I32, [P, P, P, P, P, I64, I64, I64, I32, P, P, P, P, P, P]),
    "glam_message_stack_supported": (I32, [I32, I32, I32]),
    "glam_message_stack_phase_clock": (I32, [P]),
    "glam_message_stack_bwd_supported": (I32, [I32, I32, I32, I32]),
    "glam_message_stack_bwd_workspace_bytes": (SZ, [I32, I32, I32]),
    "glam_message_stack_bwd": (I32, [P, P, P, P, P, P, P, P, P, P, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I32, I32, I32, I32,
                                     F32, I32, F32, I32, I32, F32, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "glam_message_stack_fwd": (I32, [P, P, P, I32, P, P, I32, F32, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, I64, I32, I32, I32, I32, F32, I32, F32,
                                     I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P, P, P, F32, P]),
}

_lib = None


class GlamError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library (building it is `python -m glam_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GlamError(f"{LIB_PATH} not found: build it with `python -m glam_b200.build` "
                        "(glam_b200 has no CPU or eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    from . import ABI_VERSION
    got = lib.glam_abi_version()
    if got != ABI_VERSION:
        raise GlamError(f"libglam_b200.so ABI {got} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    env = os.environ.get("GLAM_B200_MATH")
    if env:
        check(lib.glam_set_math_mode({"fp32": 0, "tf32": 1}[env.lower()]), "glam_set_math_mode")
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().glam_last_error().decode(errors="replace")
        raise GlamError(f"{what} failed (status {status}): {msg}")


def set_math_mode(mode) -> None:
    """'tf32' / 1: projections on the tcgen05 tensor cores (default); 'fp32' / 0: exact fp32 on the CUDA cores."""
    code = {"fp32": 0, "tf32": 1}.get(mode, mode)
    check(load().glam_set_math_mode(int(code)), "glam_set_math_mode")


def get_math_mode() -> str:
    return "tf32" if load().glam_get_math_mode() == 1 else "fp32"


def launch_count() -> int:
    return int(load().glam_launch_count())
