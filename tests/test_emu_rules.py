"""The warp-per-board kernel SOURCE (knightvision_b200/csrc/kv_rules.cuh), run lane-by-lane on the CPU by
tests/simt_emu, against the golden fixtures of the unmodified reference and against the oracle.  CPU only —
the same checks run on the real kernels in tests/test_gpu_rules.py."""
import numpy as np
import pytest

from knightvision_b200 import layout as L
from oracle import kv_oracle as O
import helpers as H
from simt_emu import emu


@pytest.fixture(params=[32, 16, 8], autouse=True, ids=["32_lanes_per_board", "16_lanes_per_board", "8_lanes_per_board"])
def lanes_per_board(request):
    """Every test runs the kernel source three ways: 32 lanes per board (the tree-search kernels), 16 (two boards per warp)
    and 8 (four boards per warp: what kv_rules.cu's perft / movegen / make-move kernels are built for)."""
    emu.set_width(request.param)
    yield request.param
    emu.set_width(32)


def test_emu_odd_board_counts_and_mixed_pairs():
    """Two boards per warp: a trailing odd board, pairs whose boards take different numbers of king passes / piece rounds
    (wild boards with several kings or more than 16 pieces next to tame ones), pairs with different move counts."""
    syn = H.synthetic_lines(301, 77, wild=True)
    tame = H.random_playout_positions(n_games=2, max_plies=80, seed=9)[:150]
    k = min(len(syn), len(tame))
    mix = np.empty((2 * k, 16), dtype=np.uint64)
    mix[0::2] = syn[:k]
    mix[1::2] = tame[:k]
    mix = mix[:2 * k - 1]                                   # odd count
    got = emu.movegen(mix)
    H.check_movegen_against(mix, got)
    ok = got[1] > 0
    mv = np.where(ok, got[0][np.arange(len(mix)), (got[1] - 1).clip(min=0)], 0xFFFF).astype(np.uint16)   # last legal move
    mv[::7] = 0xFFFF                                        # some boards are skipped
    out = emu.make_moves(got[3], mv)
    want = O.make_moves(got[3], mv)
    want[mv == 0xFFFF] = got[3][mv == 0xFFFF]               # 0xFFFF = leave the board untouched (kv_make_moves)
    assert np.array_equal(out[:, :13], want[:, :13])


@pytest.mark.parametrize("name,step", [("playouts", 4), ("synthetic", 1)])
def test_emu_movegen_golden(name, step):
    rows = H.load_rows(name)
    sel = slice(None, None, step)
    sub = {k: rows[k][sel] for k in ("line_in", "line_mid", "counts", "flags", "moves")}
    H.check_movegen_against(sub, emu.movegen(sub["line_in"]))


@pytest.mark.parametrize("name", ["playouts", "synthetic"])
def test_emu_make_move_golden(name):
    rows = H.load_rows(name)
    ok = rows["played"] != 0xFFFF
    out = emu.make_moves(rows["line_mid"][ok][::2], rows["played"][ok][::2])
    assert np.array_equal(out[:, :13], rows["line_out"][ok][::2][:, :13])


@pytest.mark.parametrize("name", ["playouts", "synthetic"])
def test_emu_square_under_attack_golden(name):
    rows = H.load_rows(name)
    sel = np.arange(0, len(rows["line_in"]), 5)          # the fixtures hold the 64-square mask for every 5th row
    got = emu.attacked(rows["line_in"][sel])
    assert np.array_equal(got, rows["sua"][sel])
    k = L_king_bits(rows["line_in"][sel])
    assert np.array_equal(((got >> k) & 1).astype(np.uint8), rows["incheck"][sel])


def L_king_bits(lines):
    meta = lines[:, 12]
    wtm = (meta & 1).astype(bool)
    return np.where(wtm, (meta >> 16) & 63, (meta >> 24) & 63).astype(np.uint64)


def test_emu_movegen_random_playouts_vs_oracle():
    lines = H.random_playout_positions(n_games=12, max_plies=120, seed=5)
    H.check_movegen_against(lines, emu.movegen(lines))


def test_emu_perft_vs_oracle():
    gold = H.perft_gold()
    names = list(gold)
    roots = np.array([gold[k]["line"] for k in names], dtype=np.uint64)
    for depth in (1, 2, 3):
        got = emu.perft(roots, depth)
        for i, k in enumerate(names):
            assert int(got[i, 0]) == H.PERFT_EXPECT[k][depth - 1], (k, depth)
            assert np.array_equal(got[i], O.perft2(roots[i], depth)), (k, depth)
            assert [int(x) for x in got[i, 1:5]] == gold[k]["depths"][str(depth)]["cats"]
        cnt = emu.perft(roots, -depth)            # counts-only leaves (no ordered lists, digest 0)
        assert np.array_equal(cnt[:, :5], got[:, :5]) and np.array_equal(cnt[:, 6], got[:, 6]) and not cnt[:, 5].any()


def test_emu_skip_move_and_promotion_choice_default():
    # 0xFFFF is "no move" only at the kernel level (kv_make_moves); make_move_warp itself applies any word
    line = L.start_line()[None]
    out = emu.make_moves(line, np.array([L.move_word(6, 4, 4, 4)], dtype=np.uint16))   # e2e4
    f = L.unpack_fields(out[0])
    assert f["board"][4][4] == "wp" and f["board"][6][4] == "--" and f["ep"] == (5, 4) and not f["white_to_move"]
    assert f["clock"] == 1   # pawn moves do not reset the clock (core/chessEngine.py:178)


def test_emu_synthetic_generator_lines_vs_oracle():
    """Freshly generated synthetic boards (tame and 'wild': missing / stale kings, back-rank pawns, arbitrary e.p.) through
    the kernel source on the emulator vs the pinned oracle — the CPU-sized version of the GPU fuzz test."""
    for wild, seed in ((False, 21), (True, 22)):
        syn = H.synthetic_lines(1200, seed, wild=wild)
        H.check_movegen_against(syn, emu.movegen(syn))
        assert np.array_equal(emu.attacked(syn[:200]), H.oracle_attack_masks(syn[:200]))
