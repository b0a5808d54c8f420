"""The step-wise form of the oracle's search (kvo_tree_*, used by bench.py's CPU arm to batch leaves for the fp32 network)
fed with the hash evaluator's logits must play the very games of the one-shot oracle (kvo_selfplay_game2)."""
import numpy as np

from knightvision_b200 import layout as L
from oracle import kv_oracle as O


def _play(cfg, game_id):
    t = O.Tree(cfg, L.start_line(), game_id)
    evals = 0
    while True:
        r = t.select()
        if r is False:
            break
        if r is None:
            if not t.finish_move():
                break
            continue
        line, idx = r
        lg, v = O.hash_eval(line)
        t.expand(lg[idx], v)
        evals += 1
    out = t.info()
    t.close()
    return out, evals


def test_step_api_plays_the_oracle_games():
    for sims, temp, seed, gid in ((24, 6, 5, 0), (24, 6, 5, 3), (9, 0, 2, 11)):
        cfg = O.mcts_cfg(sims, temp_plies=temp, max_plies=40, seed=seed)
        got, evals = _play(cfg, gid)
        m, lines, res = O.selfplay_game(cfg, L.start_line(), game_id=gid)
        assert got["done"] and got["ply"] == len(m) and got["result"] == res
        assert np.array_equal(got["moves"], m)
        assert 0 < evals <= sims * len(m)
