"""ctypes binding of libkv_b200.so (include/kv_b200.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KV_B200_LIB") or os.path.join(HERE, "libkv_b200.so")   # override: kernel experiments only

_lib = None

c_void_p, c_int, c_u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/kv_b200.h declares
SIGNATURES = {
    "kv_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "kv_destroy": (None, [c_void_p]),
    "kv_last_error": (ctypes.c_char_p, [c_void_p]),
    "kv_abi_version": (c_int, []),
    "kv_sm_count": (c_int, [c_void_p]),
    "kv_launch_count": (c_u64, [c_void_p]),
    "kv_profile_enable": (c_int, [c_void_p, c_int]),
    "kv_profile_read": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "kv_movegen": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "kv_make_moves": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kv_movegen_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "kv_make_moves_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "kv_attacked": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kv_attacked_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "kv_perft": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "kv_perft_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int]),
    "kv_encode": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kv_net_create": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "kv_net_set_conv_mode": (c_int, [c_void_p, c_int]),
    "kv_net_set_tower_fused": (c_int, [c_void_p, c_int]),
    "kv_net_tower_clusters4": (c_int, [c_void_p]),
    "kv_net_blob_floats": (c_u64, [c_void_p]),
    "kv_net_load": (c_int, [c_void_p, c_void_p, c_u64]),
    "kv_net_blob_device_ptr": (c_void_p, [c_void_p]),
    "kv_net_commit_weights": (c_int, [c_void_p, c_void_p]),
    "kv_net_folded_bytes": (c_u64, [c_void_p]),
    "kv_net_folded_device_ptr": (c_void_p, [c_void_p]),
    "kv_net_adopt_folded": (c_int, [c_void_p, c_void_p]),
    "kv_net_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "kv_net_forward_partial": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, ctypes.POINTER(c_int)]),
    "kv_net_forward_planes": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "kv_mcts_create": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, c_u64, c_int]),
    "kv_conv3x3_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kv_conv3x3_fprop": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "kv_conv3x3_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kv_bn_relu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_float,
                               ctypes.c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kv_bn_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kv_channel_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "kv_mcts_create_k": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, c_u64, c_int, c_int]),
    "kv_mcts_waves": (ctypes.c_int64, [c_void_p]),
    "kv_mcts_set_pipeline": (c_int, [c_void_p, c_int]),
    "kv_mcts_set_eval_split": (c_int, [c_void_p, c_int]),
    "kv_mcts_set_resign": (c_int, [c_void_p, ctypes.c_float, c_int]),
    "kv_mcts_set_root_mix": (c_int, [c_void_p, c_int]),
    "kv_mcts_set_script": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "kv_mcts_reset": (c_int, [c_void_p, c_void_p, c_u64, c_void_p]),
    "kv_mcts_run_sims": (c_int, [c_void_p, c_int, c_void_p]),
    "kv_mcts_finish_move": (c_int, [c_void_p, c_void_p]),
    "kv_mcts_run_move": (c_int, [c_void_p, c_void_p]),
    "kv_mcts_enable_cache": (c_int, [c_void_p, c_int]),
    "kv_mcts_cache_clear": (c_int, [c_void_p, c_void_p]),
    "kv_mcts_status": (c_int, [c_void_p, c_void_p, c_void_p]),
    "kv_mcts_get_roots": (c_int, [c_void_p, c_void_p, c_void_p]),
    "kv_mcts_geometry": (c_int, [c_void_p, c_void_p]),
    "kv_mcts_records": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "kv_mcts_read_root": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kv_mcts_dump_tree": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def lib():
    """Load the library (no CUDA call is made by loading it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m knightvision_b200.build` "
                "(knightvision_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class KVError(RuntimeError):
    pass


def check(ctx, rc: int, what: str):
    if rc != 0:
        msg = lib().kv_last_error(ctx)
        raise KVError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
