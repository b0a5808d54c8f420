"""ctypes wrapper of the CPU parity oracle (oracle/kv_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
Nothing under knightvision_b200/ imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkv_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libkv_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.kvo_square_under_attack.restype = ctypes.c_int
        _lib.kvo_in_check.restype = ctypes.c_int
        _lib.kvo_move_index.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


MOVE_STRIDE = 256


def movegen(lines: np.ndarray):
    """lines: uint64 [n,16] (modified in place only in the documented king-resurrection corner).
    Returns (moves u16 [n,256], counts i32 [n], flags i32 [n])."""
    lines = np.ascontiguousarray(lines, dtype=np.uint64)
    n = lines.shape[0]
    moves = np.zeros((n, MOVE_STRIDE), dtype=np.uint16)
    counts = np.zeros(n, dtype=np.int32)
    flags = np.zeros(n, dtype=np.int32)
    lib().kvo_movegen(_p(lines), ctypes.c_int(n), _p(moves), ctypes.c_int(MOVE_STRIDE), _p(counts), _p(flags))
    return moves, counts, flags, lines


def make_moves(lines: np.ndarray, mv: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(lines, dtype=np.uint64).copy()
    mv = np.ascontiguousarray(mv, dtype=np.uint16)
    lib().kvo_make_moves(_p(out), ctypes.c_int(out.shape[0]), _p(mv))
    return out


def perft(line: np.ndarray, depth: int) -> np.ndarray:
    """Returns uint64[6]: nodes, captures, ep, castles, promos, order digest."""
    line = np.ascontiguousarray(line, dtype=np.uint64)
    out = np.zeros(6, dtype=np.uint64)
    lib().kvo_perft(_p(line), ctypes.c_int(depth), _p(out))
    return out


def perft2(line: np.ndarray, depth: int) -> np.ndarray:
    """Returns uint64[8] in the layout of kv_perft (include/kv_b200.h)."""
    line = np.ascontiguousarray(line, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().kvo_perft2(_p(line), ctypes.c_int(depth), _p(out))
    return out


def square_under_attack(line: np.ndarray, r: int, c: int) -> bool:
    line = np.ascontiguousarray(line, dtype=np.uint64)
    return bool(lib().kvo_square_under_attack(_p(line), ctypes.c_int(r), ctypes.c_int(c)))


def in_check(line: np.ndarray) -> bool:
    line = np.ascontiguousarray(line, dtype=np.uint64)
    return bool(lib().kvo_in_check(_p(line)))


def encode(lines: np.ndarray) -> np.ndarray:
    lines = np.ascontiguousarray(lines, dtype=np.uint64)
    n = lines.shape[0]
    out = np.zeros((n, 12, 8, 8), dtype=np.float32)
    lib().kvo_encode(_p(lines), ctypes.c_int(n), _p(out))
    return out
