"""ctypes wrapper of the CPU lock-step emulator of the warp-per-board kernels (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libkv_emu.so")
_lib = None


def build(force: bool = False) -> str:
    deps = ([os.path.join(HERE, "kvemu.cpp")] + glob.glob(os.path.join(ROOT, "knightvision_b200", "csrc", "*.cuh"))
            + glob.glob(os.path.join(ROOT, "knightvision_b200", "csrc", "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-ffp-contract=off", "-fvisibility=hidden", "-o", SO,
                               os.path.join(HERE, "kvemu.cpp")])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_width(w: int):
    """Lanes per board of movegen / make_moves / perft below: 32 (one board per warp) or 16 (two boards per warp)."""
    lib().kvemu_set_width(ctypes.c_int(w))


def movegen(lines, stride=256):
    lines = np.ascontiguousarray(lines, dtype=np.uint64).copy()
    n = lines.shape[0]
    moves = np.zeros((n, stride), dtype=np.uint16)
    counts = np.zeros(n, dtype=np.int32)
    flags = np.zeros(n, dtype=np.int32)
    lib().kvemu_movegen(_p(lines), ctypes.c_int(n), _p(moves), ctypes.c_int(stride), _p(counts), _p(flags))
    return moves, counts, flags, lines


def attacked(lines):
    lines = np.ascontiguousarray(lines, dtype=np.uint64)
    out = np.zeros(lines.shape[0], dtype=np.uint64)
    lib().kvemu_attacked(_p(lines), ctypes.c_int(lines.shape[0]), _p(out))
    return out


def make_moves(lines, mv):
    out = np.ascontiguousarray(lines, dtype=np.uint64).copy()
    mv = np.ascontiguousarray(mv, dtype=np.uint16)
    lib().kvemu_make_moves(_p(out), ctypes.c_int(out.shape[0]), _p(mv))
    return out


def perft(roots, depth):
    roots = np.ascontiguousarray(roots, dtype=np.uint64)
    out = np.zeros((roots.shape[0], 8), dtype=np.uint64)
    lib().kvemu_perft(_p(roots), ctypes.c_int(roots.shape[0]), ctypes.c_int(depth), _p(out))
    return out


def mcts_search(start, sims, id_base=0, ply=0, edges_per_node=64, c_puct=1.5, dir_alpha=0.3, dir_eps=0.25, seed=1,
                inflight=1):
    start = np.ascontiguousarray(start, dtype=np.uint64)
    G = start.shape[0]
    moves = np.zeros((G, 256), np.uint16); N = np.zeros((G, 256), np.uint32)
    W = np.zeros((G, 256), np.float32); P = np.zeros((G, 256), np.float32); info = np.zeros((G, 4), np.int32)
    lib().kvemu_mcts_search(ctypes.c_int(G), _p(start), ctypes.c_uint64(id_base), ctypes.c_int(ply), ctypes.c_int(sims),
                            ctypes.c_int(edges_per_node), ctypes.c_float(c_puct), ctypes.c_float(dir_alpha),
                            ctypes.c_float(dir_eps), ctypes.c_uint64(seed), _p(moves), _p(N), _p(W), _p(P), _p(info),
                            ctypes.c_int(inflight))
    return moves, N, W, P, info


def set_rules(resign_thr=-0.7, resign_min_plies=15, root_mix=-1):
    """Game-loop rules of the following searches / games (defaults = kv_mcts_create_k's = the reference's)."""
    lib().kvemu_set_rules(ctypes.c_float(resign_thr), ctypes.c_int(resign_min_plies), ctypes.c_int(root_mix))


def selfplay(start, sims, max_plies, temp_plies, id_base=0, edges_per_node=64, c_puct=1.5, dir_alpha=0.3, dir_eps=0.25,
             seed=1, cache_log2=0, return_counts=False, inflight=1, pipe_order=0, script_moves=None, script_vals=None,
             return_flags=False):
    start = np.ascontiguousarray(start, dtype=np.uint64)
    G = start.shape[0]
    sm = sv = None
    if script_moves is not None or script_vals is not None:
        sm = np.ascontiguousarray(script_moves, np.uint16) if script_moves is not None else None
        sv = np.ascontiguousarray(script_vals, np.float32) if script_vals is not None else None
        stride = (sm if sm is not None else sv).shape[1]
        lib().kvemu_set_script(_p(sm) if sm is not None else None, _p(sv) if sv is not None else None, ctypes.c_int(stride))
    moves = np.zeros((G, max_plies), np.uint16); plies = np.zeros(G, np.int32); res = np.zeros(G, np.int32)
    cnt = np.zeros(2, np.int64)
    lib().kvemu_selfplay(ctypes.c_int(G), _p(start), ctypes.c_uint64(id_base), ctypes.c_int(sims),
                         ctypes.c_int(edges_per_node), ctypes.c_int(max_plies), ctypes.c_int(temp_plies),
                         ctypes.c_float(c_puct), ctypes.c_float(dir_alpha), ctypes.c_float(dir_eps),
                         ctypes.c_uint64(seed), _p(moves), _p(plies), _p(res), ctypes.c_int(cache_log2), _p(cnt),
                         ctypes.c_int(inflight), ctypes.c_int(pipe_order))
    lib().kvemu_set_script(None, None, ctypes.c_int(0))
    flags = res >> 8
    res = ((res & 0xFF) ^ 0x80) - 0x80          # sign-extend the result byte
    if return_flags:
        return moves, plies, res, flags
    if return_counts:
        return moves, plies, res, (int(cnt[0]), int(cnt[1]))
    return moves, plies, res


def tower_order(M, NT, n_layers, chunk_tiles):
    """(layer, board tile, channel tile) of every task of the whole-tower launch, in task order (kv_tower_order.h)."""
    args = (ctypes.c_int(M), ctypes.c_int(NT), ctypes.c_int(n_layers), ctypes.c_int(chunk_tiles))
    total = lib().kvemu_tower_order(*args, None)
    out = np.zeros((total, 3), np.int32)
    lib().kvemu_tower_order(*args, _p(out))
    return out


def tower_slice(part, num, den, M):
    """(first board tile, count) of a kernel's slice in the hybrid tower launch (kv_tower_order.h)."""
    out = np.zeros(2, np.int32)
    lib().kvemu_tower_slice(ctypes.c_int(part), ctypes.c_int(num), ctypes.c_int(den), ctypes.c_int(M), _p(out))
    return int(out[0]), int(out[1])
