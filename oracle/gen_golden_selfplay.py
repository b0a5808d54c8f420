#!/usr/bin/env python
"""Generate tests/golden/selfplay.json and tests/golden/gamestate_seq.json by running the UNMODIFIED reference
(/root/reference) in the build container.  Run:  python oracle/gen_golden_selfplay.py

selfplay.json — S1-S3 (scripts/self_play.py:111-255, :300-311).  `_run_single_game` is driven as it is, with
  * `_shared_model` = a scripted stub (a callable returning fixed (policy logits [B,4096], value [B,1]) per call),
  * SELFPLAY_BATCH_SIZE=1 (an environment switch of the reference, :34) so that every ply is evaluated on its own
    position instead of reusing outputs up to 15 plies old (:131-148, the buffered-inference quirk DESIGN.md §8
    does not reproduce),
  * DIR_NOISE_EPS (:12) 0 for the scripted games (the logit of the scripted move is +60, so `random.choices` (:167)
    picks it with probability 1 - 4095 e^-60) and the default 0.25 for the free-running ones,
  * for the three scenarios that cannot start from `GameState()`, the module-level name `GameState` (:87) bound to a
    factory that returns an unmodified reference GameState with its fields assigned (what the reference's own tests
    do, tests/test_castling.py:9-30).
Per game: the start line, and per ply the value the model returned, the move played (word + policy index); then
the outcome / stop reason the function logged (:239), the reward and the records it returned.
The "material" branch (:229-238) cannot be reached through the loop's four exits (no moves -> :217/:221, isDraw
-> :225, resign -> :213, max_moves -> :210); it is exercised here by calling the same expression on boards
(`'wR'.isupper()` is False and pawns score 0, so it is always 0).

gamestate_seq.json — E9 (core/chessEngine.py:193-197, :632-678): seeded move sequences with shuffles, after every
makeMove / undoMove the reference's getFEN(), positionCounts, the four end flags after getValidMoves(), isDraw().
"""
from __future__ import annotations

import json
import logging
import os
import random
import sys

os.environ["SELFPLAY_BATCH_SIZE"] = "1"
os.environ["LOG_LEVEL"] = "INFO"
os.environ.setdefault("SEED", "42")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402  (puts /root/reference/core on sys.path, imports chessEngine as CE)

CE = G.CE
L = G.L
GOLD = G.GOLD


def import_self_play():
    G.import_ai()
    import scripts.self_play as sp
    return sp


def sq(s):
    return 8 - int(s[1]), ord(s[0]) - ord("a")


def uci_index(u):
    (sr, sc), (er, ec) = sq(u[:2]), sq(u[2:4])
    return (sr * 8 + sc) * 64 + er * 8 + ec


class Stub:
    """policy/value 'model': call i returns logits peaked on script[i] (if any) over a seeded background, and values[i]."""

    def __init__(self, moves=None, values=None, default_value=0.0, seed=0, peak=60.0, bg=0.0):
        self.moves, self.values, self.default_value = moves or [], values or [], default_value
        self.rng = np.random.default_rng(seed)
        self.peak, self.bg = peak, bg
        self.calls = []

    def __call__(self, x):
        import torch
        assert x.shape[0] == 1, "SELFPLAY_BATCH_SIZE=1: one position per call"
        i = len(self.calls)
        logits = (self.rng.standard_normal(4096) * self.bg).astype(np.float32)
        if i < len(self.moves) and self.moves[i]:
            logits[uci_index(self.moves[i])] = self.peak
        v = np.float32(self.values[i] if i < len(self.values) else self.default_value)
        self.calls.append(float(v))
        return torch.from_numpy(logits)[None], torch.tensor([[v]], dtype=torch.float32)


class Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.done = None

    def emit(self, rec):
        if "complete. Moves played" in str(rec.msg):
            self.done = rec.args   # (game no, moves played, outcome, reason)


def custom_state(board_rows, white_to_move, wk, bk):
    gs = CE.GameState()
    gs.board = [row[:] for row in board_rows]
    gs.whiteToMove = white_to_move
    gs.whiteKingLocation, gs.blackKingLocation = wk, bk
    return gs


def empty_board():
    return [["--"] * 8 for _ in range(8)]


def run_case(sp, name, stub, max_moves, eps, start=None, seed=1):
    sp.EPSILON = eps
    random.seed(seed)
    np.random.seed(seed)
    if hasattr(sp._run_single_game, "_last_outputs"):
        del sp._run_single_game._last_outputs
    sp._shared_model = stub
    orig = sp.GameState
    start_gs = start() if start else CE.GameState()
    start_line = G.gs_to_line(start_gs)
    if start:
        sp.GameState = start
    cap = Capture()
    sp.logger.addHandler(cap)
    sp.logger.setLevel(logging.INFO)
    try:
        idx, recs = sp._run_single_game(0, 0.0, max_moves)
    finally:
        sp.GameState = orig
        sp.logger.removeHandler(cap)
    # replay the returned move indices on a fresh reference state to obtain the move words
    gs = start() if start else CE.GameState()
    words, lines = [], []
    for state, mi, reward in recs:
        assert isinstance(state, np.ndarray) and state.dtype == np.float32 and state.shape == (12, 8, 8)
        line = G.gs_to_line(gs)
        assert np.array_equal(state, np.stack([[[(int(line[p]) >> (r * 8 + c)) & 1 for c in range(8)] for r in range(8)]
                                               for p in range(12)]).astype(np.float32))
        mv = [m for m in gs.getValidMoves() if (m.startRow * 8 + m.startCol) * 64 + m.endRow * 8 + m.endCol == mi]
        assert len(mv) == 1, (name, mi)
        words.append(G.mv_word(mv[0]))
        lines.append([int(x) for x in line[:12]])
        gs.makeMove(mv[0])
    _, played, outcome, reason = cap.done
    assert played == len(recs)
    rewards = sorted({float(r[2]) for r in recs})
    assert len(rewards) <= 1
    out = dict(name=name, max_moves=max_moves, eps=eps, start_line=[int(x) for x in start_line],
               values=stub.calls[:len(recs)], n_model_calls=len(stub.calls), move_words=words,
               move_index=[int(r[1]) for r in recs], record_bitboards=lines, outcome=int(outcome), reason=str(reason),
               reward=rewards[0] if rewards else None, record_types=[type(recs[0][0]).__name__, str(recs[0][0].dtype),
                                                                       list(recs[0][0].shape), type(recs[0][1]).__name__,
                                                                       type(recs[0][2]).__name__] if recs else [],
               final_line=[int(x) for x in G.gs_to_line(gs)], final_white_to_move=bool(gs.whiteToMove))
    print(f"{name:28s} plies {len(recs):3d} outcome {outcome:2d} reward {out['reward']} ({reason})", flush=True)
    return out


def gen_selfplay():
    sp = import_self_play()
    assert sp.BATCH_SIZE == 1
    cases = []
    # 1. fool's mate: black mates -> outcome -1 (:217-220)
    fools = ["f2f3", "e7e5", "g2g4", "d8h4"]
    cases.append(run_case(sp, "mate_black_wins", Stub(fools), 80, 0.0))
    # 2. scholar's mate: white mates -> outcome +1
    scholars = ["e2e4", "e7e5", "d1h5", "b8c6", "f1c4", "g8f6", "h5f7"]
    cases.append(run_case(sp, "mate_white_wins", Stub(scholars), 80, 0.0))
    # 3. mate on the very move that reaches max_moves: the cap wins (:196-199, :209-211) -> draw
    cases.append(run_case(sp, "mate_at_max_moves_is_draw", Stub(fools), 4, 0.0))

    # 4. stalemate (:221-224): bK h8, wK f7, wQ g5, white plays Qg6
    def stale_start():
        b = empty_board()
        b[0][7] = "bK"; b[1][5] = "wK"; b[3][6] = "wQ"
        return custom_state(b, True, (1, 5), (0, 7))
    cases.append(run_case(sp, "stalemate", Stub(["g5g6"]), 80, 0.0, stale_start))
    # 5. stalemate on move 17 with a resigning value: the resign rule is tested first (:185-189) -> +1, not 0
    shuffle = ["g5g4", "h8h7", "g4g5", "h7h8"] * 4 + ["g5g6"]
    cases.append(run_case(sp, "resign_beats_stalemate", Stub(shuffle, [0.0] * 16 + [-0.9]), 80, 0.0, stale_start))
    # 6. same line, value exactly at the threshold (-0.7 is not < -0.7): stalemate
    cases.append(run_case(sp, "threshold_is_strict", Stub(shuffle, [0.0] * 16 + [-0.7]), 80, 0.0, stale_start))

    # 7. only kings (isDraw, :180-182): wK d4 takes the last piece
    def kings_start():
        b = empty_board()
        b[4][3] = "wK"; b[0][7] = "bK"; b[3][4] = "bp"
        return custom_state(b, True, (4, 3), (0, 7))
    cases.append(run_case(sp, "only_kings", Stub(["d4e5"]), 80, 0.0, kings_start))
    # 8. only kings and resign value on the same move: isDraw is tested first (:180) -> draw
    def kings_late_start():
        b = empty_board()
        b[0][7] = "bK"; b[4][3] = "wK"; b[3][3] = "bp"
        return custom_state(b, True, (4, 3), (0, 7))
    kshuffle = ["d4d3", "h8h7", "d3d4", "h7h8"] * 4 + ["d4d5"]
    cases.append(run_case(sp, "only_kings_beats_resign", Stub(kshuffle, [0.0] * 16 + [-0.9]), 80, 0.0, kings_late_start))
    # 9. resignation: low value from the start, triggers after move 16 (move_count > 15) with white to move -> -1
    cases.append(run_case(sp, "resign_move16", Stub(default_value=-0.95, seed=3, bg=1.0), 80, 0.25, seed=5))
    # 10. resignation first possible on move 17 (black to move afterwards) -> +1
    cases.append(run_case(sp, "resign_move17", Stub(values=[0.1] * 16, default_value=-0.8, seed=4, bg=1.0), 80, 0.25, seed=6))
    # 11. a low value before move 16 does not resign (move_count > 15 is strict): value -0.9 on plies 0..14 only
    cases.append(run_case(sp, "no_resign_before_move16", Stub(values=[-0.9] * 15, default_value=0.3, seed=5, bg=1.0), 24,
                          0.25, seed=7))
    # 12. max_moves cap (:196-199) -> draw, reward 0.2
    cases.append(run_case(sp, "max_moves_10", Stub(default_value=0.2, seed=6, bg=2.0), 10, 0.25, seed=8))
    # 13-16. free-running games, default noise, peaked random policies
    for k in range(4):
        cases.append(run_case(sp, f"free_{k}", Stub(default_value=0.05 * k, seed=10 + k, bg=3.0), 60, 0.25, seed=20 + k))
    # 17. max_moves=None (scripts/learn.py:108-109 default): no cap; ends by resignation here
    cases.append(run_case(sp, "no_cap_resigns", Stub(values=[0.0] * 40, default_value=-0.71, seed=30, bg=1.0), None, 0.25,
                          seed=31))

    # the material expression of :229-238 on a few boards (unreachable through the loop; always 0)
    mat = []
    rng = random.Random(3)
    for _ in range(8):
        gs = G.synthetic_state(rng, wild=False)
        w = sum(sp.piece_value(p) for r in gs.board for p in r if p.isupper())
        b = sum(sp.piece_value(p) for r in gs.board for p in r if p.islower())
        mat.append([w, b])
    # generate_self_play_data's filter (:300-311) on synthetic record lists
    filt = []
    for rewards in ([1.0] * 9 + [0.2] * 5, [1.0] * 6 + [-1.0] * 4 + [0.2] * 7, [0.2] * 12, []):
        recs = [(np.zeros((12, 8, 8), np.float32), i, r) for i, r in enumerate(rewards)]
        orig = sp.self_play
        sp.self_play = lambda model, n, dev, mm=None, _r=recs: list(_r)
        try:
            out = sp.generate_self_play_data(None, 0, None)
        finally:
            sp.self_play = orig
        filt.append(dict(rewards=rewards, kept=[int(r[1]) for r in out]))
    with open(os.path.join(GOLD, "selfplay.json"), "w") as f:
        json.dump(dict(cases=cases, material=mat, decisive_filter=filt,
                       resign=dict(threshold=-0.7, min_moves=15), batch_size=sp.BATCH_SIZE), f)
    print("selfplay.json written:", len(cases), "games")


# ---------------------------------------------------------------------------------------------------
def snapshot(gs, with_moves=True):
    fen = gs.getFEN()
    d = dict(fen=fen, count_cur=int(gs.positionCounts.get(fen, 0)), count_keys=len(gs.positionCounts),
             count_sum=int(sum(gs.positionCounts.values())), clock=int(gs.halfMoveClock), is_draw=bool(gs.isDraw()),
             line=[int(x) for x in G.gs_to_line(gs)])
    if with_moves:
        n = len(gs.getValidMoves())
        d.update(n_moves=n, checkMate=bool(gs.checkMate), staleMate=bool(gs.staleMate), draw50=bool(gs.draw50),
                 drawRepetition=bool(gs.drawRepetition), in_check=bool(gs.inCheck()))
    return d


def gen_gamestate_seq():
    seqs = []
    # knight shuffles from the initial position: repetition counts, the never-counted initial position (Q13)
    shuffle = ["g1f3", "g8f6", "f3g1", "f6g8"] * 3 + ["e2e4", "e7e5", "g1f3", "b8c6", "f1c4", "g8f6", "e1g1"]
    def play(gs, u):
        (sr, sc), (er, ec) = sq(u[:2]), sq(u[2:4])
        mv = [m for m in gs.getValidMoves() if (m.startRow, m.startCol, m.endRow, m.endCol) == (sr, sc, er, ec)]
        assert len(mv) == 1, u
        gs.makeMove(mv[0])
        return G.mv_word(mv[0])
    gs = CE.GameState()
    steps = [dict(op="start", **snapshot(gs))]
    for u in shuffle:
        w = play(gs, u)
        steps.append(dict(op="move", uci=u, word=w, **snapshot(gs)))
    for _ in range(3):      # undo never decrements positionCounts (Q13) and clears moved-flags (Q12)
        gs.undoMove()
        steps.append(dict(op="undo", **snapshot(gs)))
    seqs.append(dict(name="knight_shuffle", steps=steps, final_counts=dict(gs.positionCounts)))
    # seeded random playouts with undo sprinkled in
    for k in range(4):
        rng = random.Random(100 + k)
        gs = CE.GameState()
        steps = [dict(op="start", **snapshot(gs))]
        for ply in range(70):
            moves = gs.getValidMoves()
            if not moves:
                break
            # prefer reversible shuffles so repetition counts climb
            back = [m for m in moves if gs.moveLog and len(gs.moveLog) >= 2 and
                    (m.startRow, m.startCol, m.endRow, m.endCol) ==
                    (gs.moveLog[-2].endRow, gs.moveLog[-2].endCol, gs.moveLog[-2].startRow, gs.moveLog[-2].startCol)]
            m = rng.choice(back) if back and rng.random() < 0.6 else rng.choice(moves)
            gs.makeMove(m)
            steps.append(dict(op="move", uci=m.getChessNotation(), word=G.mv_word(m), **snapshot(gs)))
            if rng.random() < 0.08:
                gs.undoMove()
                steps.append(dict(op="undo", **snapshot(gs)))
        seqs.append(dict(name=f"playout_{k}", steps=steps, final_counts=dict(gs.positionCounts)))
    # loadFEN (Q14): sets board / side / e.p., not the king locations, moved flags or clock
    fens = ["r3k2r/8/8/8/8/8/8/R3K2R w - - 0 1", "4k3/2q5/8/8/1n6/8/3B4/R3K2R b KQ - 5 20",
            "8/8/8/3pP3/8/8/8/4K2k w - d6 0 1", "rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 2"]
    loads = []
    for fen in fens:
        gs = CE.GameState()
        gs.loadFEN(fen)
        d = snapshot(gs)
        d.update(fen_in=fen, board=[row[:] for row in gs.board], ep=list(gs.enPassantPossible), wk=list(gs.whiteKingLocation), bk=list(gs.blackKingLocation),
                 moves=[G.mv_word(m) for m in gs.getValidMoves()])
        loads.append(d)
    with open(os.path.join(GOLD, "gamestate_seq.json"), "w") as f:
        json.dump(dict(sequences=seqs, load_fen=loads), f)
    print("gamestate_seq.json written:", sum(len(s["steps"]) for s in seqs), "steps")


if __name__ == "__main__":
    gen_selfplay()
    gen_gamestate_seq()
