"""S1-S3 pinned to the reference: tests/golden/selfplay.json holds games the UNMODIFIED scripts/self_play.py
`_run_single_game` played with a scripted stub model (oracle/gen_golden_selfplay.py).  Replaying each game's moves and
model values through (a) the oracle's game loop, (b) the kernel source on the CPU emulator and (c) the B200 engine must
give the reference's stop ply, outcome, reward and records — mate, stalemate, only-kings, max_moves, resignation and the
orderings between them (resign beats stalemate, only-kings beats resign, the cap beats a mating move)."""
import json
import os

import numpy as np
import pytest

from knightvision_b200 import layout as L
from oracle import kv_oracle as O
from simt_emu import emu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selfplay.json")
NO_CAP = 200


def cases():
    with open(GOLD) as f:
        return json.load(f)


def _script(c, stride):
    mv = np.full(stride, 0xFFFF, np.uint16)
    va = np.full(stride, np.nan, np.float32)
    n = len(c["move_words"])
    mv[:n] = c["move_words"]
    va[:n] = c["values"]
    return mv, va


def _reward(result):
    return 1.0 if result > 0 else (-1.0 if result < 0 else 0.2)     # scripts/self_play.py:245-250


def _check(c, moves, lines, result):
    assert len(moves) == len(c["move_words"]), c["name"]
    assert [int(m) for m in moves] == c["move_words"], c["name"]
    assert [O.lib().kvo_move_index(int(m)) for m in moves] == c["move_index"], c["name"]
    assert np.array_equal(np.asarray(lines, np.uint64)[:, :12], np.array(c["record_bitboards"], np.uint64)), c["name"]
    assert result == c["outcome"], (c["name"], c["reason"])
    assert _reward(result) == c["reward"], c["name"]


def test_fixture_covers_every_exit_of_the_reference_loop():
    g = cases()
    reasons = {c["reason"].split(" (")[0].split(" ")[0] for c in g["cases"]}
    assert {"Checkmate", "Stalemate", "Resignation", "Max", "Draw"} <= reasons
    assert g["resign"] == {"threshold": -0.7, "min_moves": 15} and g["batch_size"] == 1
    for c in g["cases"]:
        assert c["record_types"] == ["ndarray", "float32", [12, 8, 8], "int", "float"]
    assert all(w == 0 and b == 0 for w, b in g["material"])      # :229-238 scores nothing: always a draw


def test_oracle_game_loop_matches_reference_games():
    for c in cases()["cases"]:
        cap = c["max_moves"] or NO_CAP
        cfg = O.mcts_cfg(1, max_plies=cap, seed=3)
        mv, va = _script(c, cap)
        moves, lines, result, flags = O.selfplay_game(cfg, np.array(c["start_line"], np.uint64), script_moves=mv,
                                                      script_vals=va, return_flags=True)
        assert flags == 0
        _check(c, moves, lines, result)


def test_kernel_source_game_loop_matches_reference_games():
    """mcts_finish_move_warp (the CUDA source, on the lock-step emulator) over the same scripts."""
    cs = cases()["cases"]
    for cap in sorted({c["max_moves"] or NO_CAP for c in cs}):
        grp = [c for c in cs if (c["max_moves"] or NO_CAP) == cap]
        start = np.array([c["start_line"] for c in grp], np.uint64)
        sm = np.stack([_script(c, cap)[0] for c in grp])
        sv = np.stack([_script(c, cap)[1] for c in grp])
        moves, plies, res, flags = emu.selfplay(start, sims=1, max_plies=cap, temp_plies=0, seed=3, script_moves=sm,
                                                script_vals=sv, return_flags=True)
        for i, c in enumerate(grp):
            assert flags[i] == 0 and plies[i] == len(c["move_words"]), c["name"]
            assert [int(m) for m in moves[i, :plies[i]]] == c["move_words"], c["name"]
            assert int(res[i]) == c["outcome"], (c["name"], c["reason"])


def test_resign_rule_can_be_switched_off_and_illegal_script_moves_are_flagged():
    c = next(x for x in cases()["cases"] if x["name"] == "resign_move16")
    cfg = O.mcts_cfg(1, max_plies=20, seed=3, resign_min_plies=-1)
    mv, va = _script(c, 20)
    moves, lines, result = O.selfplay_game(cfg, np.array(c["start_line"], np.uint64), script_moves=mv, script_vals=va)
    assert len(moves) == 20 and result == 0                      # played on to the cap
    bad = mv.copy()
    bad[2] = L.move_word(0, 0, 4, 4, False, False, False)        # a8e4 is never legal on move 3
    moves, lines, result, flags = O.selfplay_game(O.mcts_cfg(1, max_plies=20, seed=3), np.array(c["start_line"], np.uint64),
                                                  script_moves=bad, script_vals=va, return_flags=True)
    assert flags == 2 and len(moves) == 2
    emu_m, emu_p, emu_r, emu_f = emu.selfplay(np.array([c["start_line"]], np.uint64), sims=1, max_plies=20, temp_plies=0,
                                              seed=3, script_moves=bad[None], script_vals=va[None], return_flags=True)
    assert emu_f[0] == 2 and emu_p[0] == 2


def test_decisive_filter_matches_reference():
    from knightvision_b200.selfplay import filter_decisive
    for f in cases()["decisive_filter"]:
        recs = [(None, i, r) for i, r in enumerate(f["rewards"])]
        assert [r[1] for r in filter_decisive(recs)] == f["kept"]


@pytest.mark.gpu
def test_gpu_game_loop_matches_reference_games():
    import torch
    from knightvision_b200.engine import Engine, lines_to_device
    eng = Engine(0)
    cs = cases()["cases"]
    for cap in sorted({c["max_moves"] or NO_CAP for c in cs}):
        grp = [c for c in cs if (c["max_moves"] or NO_CAP) == cap]
        G = len(grp)
        eng.mcts_create(G, 1, max_plies=cap, temp_plies=0, seed=3, eval_mode=0)
        sm = torch.from_numpy(np.stack([_script(c, cap)[0] for c in grp]).view(np.int16)).to(eng.device)
        sv = torch.from_numpy(np.stack([_script(c, cap)[1] for c in grp])).to(eng.device)
        eng.mcts_set_script(sm, sv)
        eng.mcts_reset(lines_to_device(np.array([c["start_line"] for c in grp], np.uint64), eng.device), 0)
        for _ in range(cap + 1):
            eng.mcts_run_move()
        st = eng.mcts_status()
        assert st["done"] == G and st["script_misses"] == 0 and st["overflow"] == 0
        lines, move, reward, game = (t.cpu().numpy() for t in eng.mcts_records())
        lines = lines.view(np.uint64)
        for i, c in enumerate(grp):
            sel = game == i
            assert int(sel.sum()) == len(c["move_index"]), c["name"]
            assert move[sel].tolist() == c["move_index"], c["name"]
            assert np.array_equal(lines[sel][:, :12], np.array(c["record_bitboards"], np.uint64)), c["name"]
            assert np.allclose(reward[sel], c["reward"]), (c["name"], c["reason"])
            # and the reference's float planes (record format, scripts/self_play.py:173-174)
            planes = eng.encode(torch.from_numpy(lines[sel].view(np.int64)).to(eng.device)).cpu().numpy()
            assert planes.dtype == np.float32 and planes.shape[1:] == (12, 8, 8)
            assert np.array_equal(planes, O.encode(lines[sel]))
        eng.mcts_set_script(None, None)
    eng.close()


def test_decisive_filter_on_packed_device_records_matches_reference():
    """learn.reinforcement_loop filters the packed records (scripts/learn.py:186 -> generate_self_play_data): same rule."""
    import torch
    from knightvision_b200.selfplay import filter_decisive_device
    for f in cases()["decisive_filter"]:
        n = len(f["rewards"])
        lines = torch.arange(n * 16, dtype=torch.int64).reshape(n, 16)
        move = torch.arange(n, dtype=torch.int32)
        reward = torch.tensor(f["rewards"], dtype=torch.float32)
        l2, m2, r2 = filter_decisive_device(lines, move, reward)
        assert m2.tolist() == f["kept"] and l2[:, 0].tolist() == [16 * k for k in f["kept"]]
        assert r2.tolist() == [pytest.approx(f["rewards"][k]) for k in f["kept"]]
