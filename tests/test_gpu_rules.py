"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the golden fixtures of the
unmodified reference and against the oracle.  Bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from knightvision_b200 import layout as L
from oracle import kv_oracle as O
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from knightvision_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _movegen(eng, lines):
    from knightvision_b200.engine import lines_to_device, lines_to_host
    d = lines_to_device(lines, eng.device)
    moves, counts, flags = eng.movegen(d)
    torch.cuda.synchronize()
    return (moves.cpu().numpy().view(np.uint16), counts.cpu().numpy(), flags.cpu().numpy(), lines_to_host(d))


@pytest.mark.parametrize("name", ["playouts", "synthetic"])
def test_movegen_golden(eng, name):
    rows = H.load_rows(name)
    H.check_movegen_against(rows, _movegen(eng, rows["line_in"]))
    H.check_movegen_against(rows, eng.movegen_host(rows["line_in"]))   # host-buffer C-ABI form


@pytest.mark.parametrize("name", ["playouts", "synthetic"])
def test_make_move_golden(eng, name):
    from knightvision_b200.engine import lines_to_device, lines_to_host
    rows = H.load_rows(name)
    d = lines_to_device(rows["line_mid"], eng.device)
    mv = torch.from_numpy(rows["played"].view(np.int16)).to(eng.device)
    eng.make_moves(d, mv)
    assert np.array_equal(lines_to_host(d)[:, :13], rows["line_out"][:, :13])
    ok = rows["played"] != 0xFFFF
    out = eng.make_moves_host(rows["line_mid"][ok], rows["played"][ok])
    assert np.array_equal(out[:, :13], rows["line_out"][ok][:, :13])


@pytest.mark.parametrize("name", ["playouts", "synthetic"])
def test_square_under_attack_and_in_check_golden(eng, name):
    rows = H.load_rows(name)
    sel = np.arange(0, len(rows["line_in"]), 5)
    got = eng.attacked_host(rows["line_in"][sel])
    assert np.array_equal(got, rows["sua"][sel])
    meta = rows["line_in"][sel][:, 12]
    k = np.where((meta & 1).astype(bool), (meta >> 16) & 63, (meta >> 24) & 63).astype(np.uint64)
    assert np.array_equal(((got >> k) & 1).astype(np.uint8), rows["incheck"][sel])


def test_movegen_random_playouts_vs_oracle(eng):
    lines = H.random_playout_positions(n_games=256, max_plies=200, seed=11)
    assert len(lines) > 20000
    H.check_movegen_against(lines, _movegen(eng, lines))


def test_fuzz_large_playouts_and_wild_boards_vs_oracle(eng):
    """Scale check of the rules kernels against the pinned oracle: ~400 k playout positions (move lists in order, flags,
    state rewrite, make-move of a random legal move) and 60 000 synthetic boards, half of them 'wild' (missing / stale
    kings, back-rank pawns, arbitrary e.p. squares — the corners the reference-made synthetic fixtures probe)."""
    lines = H.random_playout_positions(n_games=2048, max_plies=260, seed=2024)
    assert len(lines) > 300000
    got = _movegen(eng, lines)
    H.check_movegen_against(lines, got)
    moves, counts = got[0], got[1]
    rng = np.random.default_rng(3)
    idx = (rng.random(len(lines)) * np.maximum(counts, 1)).astype(np.int64)
    pick = np.where(counts > 0, moves[np.arange(len(lines)), idx], 0xFFFF).astype(np.uint16)
    after = eng.make_moves_host(got[3], pick)
    live = counts > 0
    assert np.array_equal(after[live][:, :13], O.make_moves(got[3][live].copy(), pick[live])[:, :13])
    assert np.array_equal(after[~live], got[3][~live])          # 0xFFFF = leave the board untouched (kv_b200.h)
    for wild, seed in ((False, 5), (True, 6)):
        syn = H.synthetic_lines(30000, seed, wild=wild)
        g2 = _movegen(eng, syn)
        H.check_movegen_against(syn, g2)
        assert np.array_equal(eng.attacked_host(syn[:1500]), H.oracle_attack_masks(syn[:1500]))


def test_perft_golden_all_positions(eng):
    gold = H.perft_gold()
    names = list(gold)
    roots = np.array([gold[k]["line"] for k in names], dtype=np.uint64)
    for depth in (1, 2, 3, 4):
        got = eng.perft_host(roots, depth)
        for i, k in enumerate(names):
            g = gold[k]["depths"][str(depth)]
            assert int(got[i, 0]) == g["nodes"] == H.PERFT_EXPECT[k][depth - 1], (k, depth)
            assert [int(x) for x in got[i, 1:5]] == g["cats"], (k, depth)
            if depth <= 3:
                assert np.array_equal(got[i], O.perft2(roots[i], depth)), (k, depth)
        cnt = eng.perft_host(roots, depth, chunk=-65536)     # counts-only leaves
        assert np.array_equal(cnt[:, :5], got[:, :5]) and np.array_equal(cnt[:, 6], got[:, 6]) and not cnt[:, 5].any()


def test_perft5_startpos_is_the_custom_count(eng):
    got = eng.perft_host(L.start_line()[None], 5)[0]
    g = H.perft_gold()["startpos"]["depths"]["5"]
    assert int(got[0]) == 4865721 == g["nodes"]          # standard chess: 4 865 609
    assert [int(x) for x in got[1:5]] == g["cats"]
    assert np.array_equal(got, O.perft2(L.start_line(), 5))   # digest + movegen-call count


def test_perft_divide_matches_reference(eng):
    g = H.perft_gold()["startpos"]
    from knightvision_b200.engine import Engine  # noqa: F401
    kids = O.make_moves(np.repeat(L.start_line()[None], 20, 0), np.array(g["root_moves"], dtype=np.uint16))
    got = eng.perft_host(kids, 4)
    assert [int(x) for x in got[:, 0]] == g["depths"]["5"]["divide"]


def test_full_batch_65536_boards(eng):
    """BASELINE config 2 size: 65 536 boards in one launch; size-independent properties."""
    from knightvision_b200.engine import lines_to_device
    gold = H.perft_gold()
    base = np.array([gold[k]["line"] for k in gold], dtype=np.uint64)
    lines = base[np.arange(65536) % len(base)]
    moves, counts, flags, after = _movegen(eng, lines)
    em, ec, ef, _ = O.movegen(base.copy())
    assert np.array_equal(counts, ec[np.arange(65536) % len(base)])
    for i in list(range(7)) + [65535 - k for k in range(7)]:
        c = ec[i % len(base)]
        assert np.array_equal(moves[i, :c], em[i % len(base), :c])
    assert np.array_equal(after, lines)
    # a perft of perfts: sum over the batch equals count * per-position value (depth 2, chunked launches)
    d = lines_to_device(lines, eng.device)
    out = eng.perft(d, 2).cpu().numpy().view(np.uint64)
    exp = np.array([H.PERFT_EXPECT[k][1] for k in gold], dtype=np.uint64)
    assert np.array_equal(out[:, 0], exp[np.arange(65536) % len(base)])
    assert np.array_equal(out[:, 6], (1 + ec)[np.arange(65536) % len(base)].astype(np.uint64))


def test_empty_and_single(eng):
    assert eng.movegen_host(np.zeros((0, 16), dtype=np.uint64))[1].shape == (0,)
    m, c, f, _ = eng.movegen_host(L.start_line()[None])
    assert c[0] == 20 and f[0] == 0
    assert [L.move_uci(int(x)) for x in m[0, :20]][:4] == ["a2a3", "a2a4", "b2b3", "b2b4"]


def test_encode_matches_reference(eng):
    from knightvision_b200.engine import lines_to_device
    g = np.load(H.GOLDEN + "/encode.npz")
    out = eng.encode(lines_to_device(g["lines"], eng.device)).cpu().numpy()
    assert out.dtype == np.float32 and np.array_equal(out, g["planes"])
