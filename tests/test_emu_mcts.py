"""The PUCT tree-search kernel SOURCE (kv_mcts.cuh) run on the CPU emulator against the sequential MCTS oracle
(oracle/kv_oracle.c): visit counts, W sums, priors and whole games must be bit-exact.  CPU only."""
import numpy as np

import helpers as H
from knightvision_b200 import layout as L
from oracle import kv_oracle as O
from simt_emu import emu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_detmath_is_close_to_libm():
    # the deterministic exp/log are ordinary single-precision approximations (sanity, not parity)
    import ctypes
    xs = np.linspace(-20, 5, 101, dtype=np.float32)
    src = r'''
    #include "include/kv_detmath.h"
    float e(float x){return kvd_expf(x);} float l(float x){return kvd_logf(x);}
    float g(float a, unsigned long long s, unsigned long long i){return kvd_gamma_small(a, s, i, 0);}
    '''
    import os, subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        open(d + "/t.c", "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", root, "-o", d + "/t.so",
                               d + "/t.c", "-lm"])
        t = ctypes.CDLL(d + "/t.so")
        t.e.restype = t.l.restype = t.g.restype = ctypes.c_float
        t.e.argtypes = t.l.argtypes = [ctypes.c_float]
        t.g.argtypes = [ctypes.c_float, ctypes.c_uint64, ctypes.c_uint64]
        for x in xs:
            assert abs(t.e(float(x)) - np.exp(np.float64(x))) <= 2e-6 * np.exp(np.float64(x)) + 1e-30
        for x in np.geomspace(1e-6, 1e6, 61):
            assert abs(t.l(float(x)) - np.log(x)) < 2e-6 * max(1.0, abs(np.log(x)))
        gs = np.array([t.g(0.3, 11, i) for i in range(4000)])
        assert abs(gs.mean() - 0.3) < 0.03 and (gs > 0).all()     # Gamma(0.3, 1): mean 0.3


def test_emu_search_matches_oracle():
    lines = H.random_playout_positions(n_games=3, max_plies=60, seed=21)
    roots = np.concatenate([L.start_line()[None], lines[[7, 40, 90, 130]]])
    for sims, seed in ((1, 3), (40, 9)):
        mv, N, W, P, info = emu.mcts_search(roots, sims=sims, id_base=17, ply=2, seed=seed)
        for g in range(len(roots)):
            r = O.mcts_search(O.mcts_cfg(sims, seed=seed), roots[g], game_id=17 + g, ply=2)
            n = int(info[g, 0])
            assert n == len(r["moves"]) and info[g, 1] == r["nodes"] and info[g, 2] == r["edges"]
            assert np.array_equal(mv[g, :n], r["moves"]) and np.array_equal(N[g, :n], r["N"])
            assert np.array_equal(_bits(W[g, :n]), _bits(r["W"])) and np.array_equal(_bits(P[g, :n]), _bits(r["P"]))
            if sims > 1:
                assert int(r["N"].sum()) == sims - 1


def test_emu_selfplay_games_match_oracle():
    start = np.stack([L.start_line()] * 2)
    moves, plies, res = emu.selfplay(start, sims=12, max_plies=24, temp_plies=6, id_base=40, seed=5)
    for g in range(2):
        cfg = O.mcts_cfg(12, temp_plies=6, max_plies=24, seed=5)
        m, lines, r = O.selfplay_game(cfg, start[g], game_id=40 + g)
        assert len(m) == plies[g] and r == res[g] and np.array_equal(m, moves[g, :plies[g]])


def test_edge_pool_overflow_rule_is_shared():
    # edges_per_node = 1 forces the "expansion does not fit -> terminal draw" rule on both sides
    roots = L.start_line()[None]
    mv, N, W, P, info = emu.mcts_search(roots, sims=300, edges_per_node=1, seed=2)
    cfg = O.mcts_cfg(300, edges_per_node=1, seed=2)
    r = O.mcts_search(cfg, roots[0])
    assert r["overflow"] == 1 and info[0, 3] == 1
    n = int(info[0, 0])
    assert np.array_equal(N[0, :n], r["N"]) and np.array_equal(_bits(W[0, :n]), _bits(r["W"]))


def test_eval_cache_does_not_change_games():
    """Cache hits / in-wave followers re-derive the same priors and values: identical games, fewer evaluations."""
    start = np.stack([L.start_line()] * 6)
    base = emu.selfplay(start, sims=16, max_plies=10, temp_plies=4, id_base=7, seed=11, return_counts=True)
    for log2 in (8, 14):    # tiny table (evictions, window pressure) and roomy table
        got = emu.selfplay(start, sims=16, max_plies=10, temp_plies=4, id_base=7, seed=11, cache_log2=log2,
                           return_counts=True)
        assert np.array_equal(base[0], got[0]) and np.array_equal(base[1], got[1]) and np.array_equal(base[2], got[2])
        evals, late = got[3]
        assert late > 0 and evals + late == base[3][0]      # every expansion is either a tower eval or a cache serve
    assert base[3][1] == 0


def test_pipelined_groups_do_not_change_games():
    """Two game groups with alternating waves (kv_mcts.cu mcts_run_waves): every interleaving the stream/event graph
    admits gives the games of the plain schedule, with and without the shared evaluation cache (cross-group followers
    included), for K = 1 and with virtual loss."""
    lines = H.random_playout_positions(n_games=2, max_plies=40, seed=12)
    # 20 games = groups of 16 + 4; the initial position is in both groups, so leaves of one group follow leaders of the other
    start = np.concatenate([np.stack([L.start_line()] * 8), lines[[3, 9, 17, 25, 31, 38]], np.stack([L.start_line()] * 6)])
    for K, sims, cases in ((1, 12, ((0, 2), (7, 1), (7, 2), (7, 3), (13, 3))), (4, 16, ((7, 2), (13, 3)))):
        base = emu.selfplay(start, sims=sims, max_plies=4, temp_plies=2, id_base=21, seed=13, inflight=K,
                            return_counts=True)
        for log2, order in cases:
            got = emu.selfplay(start, sims=sims, max_plies=4, temp_plies=2, id_base=21, seed=13, inflight=K,
                               cache_log2=log2, return_counts=True, pipe_order=order)
            assert np.array_equal(base[0], got[0]) and np.array_equal(base[1], got[1])
            assert np.array_equal(base[2], got[2])
            evals, late = got[3]
            assert evals + late == base[3][0]
            assert (late > 0) == (log2 > 0)


def test_policy_sampling_mode_sims_1():
    """sims = 1 is the reference's move rule (sample from softmax + Dirichlet over the legal moves, no search)."""
    start = np.stack([L.start_line()] * 3)
    moves, plies, res = emu.selfplay(start, sims=1, max_plies=30, temp_plies=30, id_base=3, seed=8)
    seen = set()
    for g in range(3):
        m, lines, r = O.selfplay_game(O.mcts_cfg(1, temp_plies=30, max_plies=30, seed=8), start[g], game_id=3 + g)
        assert np.array_equal(m, moves[g, :plies[g]]) and r == res[g]
        seen.add(int(m[0]))
    assert len(seen) > 1          # not always the first legal move


def test_virtual_loss_inflight_search_matches_oracle():
    """K simulations in flight per game and wave (virtual loss): the kernel source on the emulator
    against the oracle's wave emulation, bit for bit; K = 1 is the sequential search."""
    lines = H.random_playout_positions(n_games=3, max_plies=60, seed=33)
    roots = np.concatenate([L.start_line()[None], lines[[5, 30, 77, 120]]])
    seq = emu.mcts_search(roots, sims=96, id_base=5, ply=1, seed=4)
    for K in (1, 2, 4, 8, 16):
        mv, N, W, P, info = emu.mcts_search(roots, sims=96, id_base=5, ply=1, seed=4, inflight=K)
        if K == 1:
            assert np.array_equal(N, seq[1]) and np.array_equal(_bits(W), _bits(seq[2]))
        differs = False
        for g in range(len(roots)):
            r = O.mcts_search(O.mcts_cfg(96, seed=4, inflight=K), roots[g], game_id=5 + g, ply=1)
            n = int(info[g, 0])
            assert n == len(r["moves"]) and info[g, 1] == r["nodes"] and info[g, 2] == r["edges"]
            assert np.array_equal(mv[g, :n], r["moves"]) and np.array_equal(N[g, :n], r["N"])
            assert np.array_equal(_bits(W[g, :n]), _bits(r["W"])) and np.array_equal(_bits(P[g, :n]), _bits(r["P"]))
            assert int(r["N"].sum()) == 96 - 1                    # exactly `sims` simulations, whatever the wave shape
            assert int(N[g, :n].max()) < (1 << 24)                # no virtual visit left behind in the counters
            differs |= not np.array_equal(N[g, :n], seq[1][g, :n])
        if K >= 4:
            assert differs          # virtual loss does change the search (it is not silently sequential)


def test_virtual_loss_selfplay_games_match_oracle():
    start = np.stack([L.start_line()] * 3)
    for K, cache in ((4, 0), (8, 10)):
        moves, plies, res = emu.selfplay(start, sims=24, max_plies=16, temp_plies=5, id_base=9, seed=6, inflight=K,
                                         cache_log2=cache)
        for g in range(3):
            cfg = O.mcts_cfg(24, temp_plies=5, max_plies=16, seed=6, inflight=K)
            m, lines, r = O.selfplay_game(cfg, start[g], game_id=9 + g)
            assert len(m) == plies[g] and r == res[g] and np.array_equal(m, moves[g, :plies[g]])


def _reference_rule_priors(line, moves, alpha, eps, seed, game_id, ply):
    """scripts/self_play.py:150-167 in float64: softmax over all 4096 logits, Dirichlet noise over all 4096 indices,
    (1-eps) p + eps noise, legal entries renormalised."""
    lg, _ = O.hash_eval(line)
    p = np.exp(lg.astype(np.float64) - lg.max())
    p /= p.sum()
    if eps > 0:
        g = O.root_noise(alpha, seed, game_id, ply).astype(np.float64)
        p = (1 - eps) * p + eps * g / g.sum()
    idx = [O.lib().kvo_move_index(int(m)) for m in moves]
    w = p[idx]
    return w / w.sum()


def test_reference_rule_mode_priors_and_games():
    """sims = 1 (no search): root priors mixed the reference's way, whole games bit-exact between the kernel source and
    the oracle, resignation included (hash values below -0.7 are frequent)."""
    lines = H.random_playout_positions(n_games=2, max_plies=40, seed=5)
    roots = np.concatenate([L.start_line()[None], lines[[11, 30, 55]]])
    for eps in (0.25, 0.0):
        mv, N, W, P, info = emu.mcts_search(roots, sims=1, id_base=3, ply=7, seed=12, dir_eps=eps)
        for g in range(len(roots)):
            n = int(info[g, 0])
            want = _reference_rule_priors(roots[g], mv[g, :n], 0.3, eps, 12, 3 + g, 7)
            assert np.allclose(P[g, :n], want, rtol=2e-5, atol=1e-7), (g, eps)
            assert abs(float(P[g, :n].sum()) - 1.0) < 1e-5
    # the legal-only mixing (search default) is a different distribution: the switch matters
    emu.set_rules(root_mix=0)
    try:
        _, _, _, P0, info0 = emu.mcts_search(roots[:1], sims=1, id_base=3, ply=7, seed=12)
    finally:
        emu.set_rules()
    assert not np.allclose(P0[0, :20], P[0, :20], rtol=1e-3)
    start = np.stack([L.start_line()] * 3)
    moves, plies, res = emu.selfplay(start, sims=1, max_plies=40, temp_plies=0, id_base=9, seed=4)
    resigned = 0
    for g in range(3):
        m, pos, r = O.selfplay_game(O.mcts_cfg(1, max_plies=40, seed=4), start[g], game_id=9 + g)
        assert len(m) == plies[g] and r == res[g] and np.array_equal(m, moves[g, :plies[g]])
        resigned += int(plies[g] < 40 and r != 0)
    assert resigned >= 1
