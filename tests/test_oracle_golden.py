"""Pins the C oracle (oracle/kv_oracle.c) against fixtures produced by the UNMODIFIED Python reference
(oracle/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from knightvision_b200 import layout as L
from oracle import kv_oracle as O


@pytest.fixture(scope="module", params=["playouts", "synthetic"])
def rows(request, golden_dir):
    return np.load(os.path.join(golden_dir, request.param + ".npz"))


def test_move_lists_in_reference_order(rows):
    moves, counts, flags, mid = O.movegen(rows["line_in"].copy())
    assert np.array_equal(counts, rows["counts"])
    nm = rows["moves"].shape[1]
    assert np.array_equal(moves[:, :nm], rows["moves"])
    assert not moves[:, nm:].any()
    assert np.array_equal(flags, rows["flags"])
    # getKingMoves' restore quirk (core/chessEngine.py:564) may rewrite the board
    assert np.array_equal(mid, rows["line_mid"])


def test_make_move(rows):
    ok = rows["played"] != 0xFFFF
    out = O.make_moves(rows["line_mid"][ok], rows["played"][ok])
    assert np.array_equal(out, rows["line_out"][ok])


def test_square_under_attack_and_in_check(rows):
    lines = rows["line_in"]
    for i in range(0, len(lines), 5):
        m = 0
        for sq in range(64):
            if O.square_under_attack(lines[i], sq >> 3, sq & 7):
                m |= 1 << sq
        assert m == int(rows["sua"][i]), i
    ic = np.array([O.in_check(l) for l in lines], dtype=np.uint8)
    assert np.array_equal(ic, rows["incheck"])


def test_fixture_coverage(golden_dir):
    """The fixtures really contain the quirky cases (king captured, mutation, e.p., castle, promo)."""
    p = np.load(os.path.join(golden_dir, "playouts.npz"))
    s = np.load(os.path.join(golden_dir, "synthetic.npz"))
    allm = np.concatenate([p["moves"].ravel(), s["moves"].ravel()])
    assert ((allm >> 12) & 1).sum() > 10      # e.p.
    assert ((allm >> 13) & 1).sum() > 10      # castle
    assert ((allm >> 14) & 1).sum() > 10      # promotion
    assert ((s["flags"] & L.RF_STATE_MUTATED) != 0).sum() > 10
    assert ((p["flags"] & L.RF_CHECKMATE) != 0).sum() + ((s["flags"] & L.RF_CHECKMATE) != 0).sum() > 3
    kingless = ((p["line_in"][:, 0] == 0) | (p["line_in"][:, 6] == 0)).sum()
    assert kingless > 0                       # a king was captured in a playout


def test_perft_counts_categories_and_order(golden_dir):
    with open(os.path.join(golden_dir, "perft.json")) as f:
        gold = json.load(f)
    expect = {"startpos": [20, 400, 8902, 197281, 4865721], "castle_w": [26, 568, 13744, 314346],
              "castle_b": [26, 568, 13744, 314346], "ep_a": [6, 38, 257, 1971], "ep_b": [6, 38, 257, 1971],
              "promo_w": [24, 462, 12448, 274548], "promo_b": [24, 462, 12448, 272623]}
    for name, e in gold.items():
        line = np.array(e["line"], dtype=np.uint64)
        for d, node_count in enumerate(expect[name], start=1):
            if d == 5:
                continue  # exercised by the divide below
            g = e["depths"][str(d)]
            assert g["nodes"] == node_count
            out = O.perft(line, d)
            assert int(out[0]) == g["nodes"], (name, d)
            assert [int(x) for x in out[1:5]] == g["cats"], (name, d)
        moves, counts, _, _ = O.movegen(line[None].copy())
        assert [int(x) for x in moves[0, :counts[0]]] == e["root_moves"]
        # divide + order digests at the deepest depth (startpos d5 = 4 865 721, not the standard 4 865 609)
        dmax = str(len(expect[name]))
        g = e["depths"][dmax]
        assert g["nodes"] == expect[name][-1]
        children = O.make_moves(np.repeat(line[None], len(e["root_moves"]), 0),
                                np.array(e["root_moves"], dtype=np.uint16))
        for k, ch in enumerate(children):
            out = O.perft(ch, int(dmax) - 1)
            assert int(out[0]) == g["divide"][k], (name, k)
            assert int(out[5]) == g["child_digests"][k], (name, k)


def test_startpos_root_order():
    moves, counts, _, _ = O.movegen(L.start_line()[None].copy())
    uci = [L.move_uci(int(m)) for m in moves[0, :counts[0]]]
    assert uci == ("a2a3 a2a4 b2b3 b2b4 c2c3 c2c4 d2d3 d2d4 e2e3 e2e4 f2f3 f2f4 g2g3 g2g4 h2h3 h2h4 "
                   "b1a3 b1c3 g1f3 g1h3").split()


def test_encode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "encode.npz"))
    assert np.array_equal(O.encode(g["lines"]), g["planes"])
    idx = [O.lib().kvo_move_index(int(w)) for w in g["move_words"]]
    assert idx == [int(x) for x in g["move_index"]]
    for i, sr, sc, er, ec in g["decode"]:
        assert (i // 64 // 8, i // 64 % 8, i % 64 // 8, i % 64 % 8) == (sr, sc, er, ec)
