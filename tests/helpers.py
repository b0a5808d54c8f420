"""Shared helpers for the parity tests (oracle-driven position generators)."""
from __future__ import annotations

import json
import os

import numpy as np

from knightvision_b200 import layout as L
from oracle import kv_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PERFT_EXPECT = {"startpos": [20, 400, 8902, 197281, 4865721], "castle_w": [26, 568, 13744, 314346],
                "castle_b": [26, 568, 13744, 314346], "ep_a": [6, 38, 257, 1971], "ep_b": [6, 38, 257, 1971],
                "promo_w": [24, 462, 12448, 274548], "promo_b": [24, 462, 12448, 272623]}


def load_rows(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def perft_gold():
    with open(os.path.join(GOLDEN, "perft.json")) as f:
        return json.load(f)


def random_playout_positions(n_games: int, max_plies: int, seed: int) -> np.ndarray:
    """Positions visited by uniformly random playouts driven by the ORACLE (already pinned to the reference)."""
    rng = np.random.default_rng(seed)
    lines = np.stack([L.start_line()] * n_games)
    alive = np.ones(n_games, dtype=bool)
    out = []
    for _ in range(max_plies):
        if not alive.any():
            break
        cur = lines[alive]
        out.append(cur.copy())
        moves, counts, flags, mid = O.movegen(cur.copy())
        pick = np.full(len(cur), 0xFFFF, dtype=np.uint16)
        ok = counts > 0
        idx = (rng.random(len(cur)) * np.maximum(counts, 1)).astype(np.int64)
        pick[ok] = moves[np.arange(len(cur)), idx][ok]
        nxt = O.make_moves(mid, pick)
        lines[alive] = nxt
        a = alive.copy()
        a[alive] = ok & ((flags & L.RF_ONLY_KINGS) == 0)
        alive = a
    return np.concatenate(out)


def check_movegen_against(rows_or_lines, got):
    """got = (moves, counts, flags, lines_after).  Compares with the golden rows (or the oracle for raw lines)."""
    moves, counts, flags, after = got
    if isinstance(rows_or_lines, np.ndarray):
        em, ec, ef, ea = O.movegen(rows_or_lines.copy())
    else:
        r = rows_or_lines
        ec, ef, ea = r["counts"], r["flags"], r["line_mid"]
        em = np.zeros((len(ec), 256), dtype=np.uint16)
        em[:, :r["moves"].shape[1]] = r["moves"]
    assert np.array_equal(counts, ec)
    assert np.array_equal(flags, ef)
    k = np.arange(256)[None, :] < np.minimum(ec, 256)[:, None]
    assert np.array_equal(np.where(k, moves[:, :256], 0), np.where(k, em, 0))
    assert np.array_equal(after[:, :13], ea[:, :13])
