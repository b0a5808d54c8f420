"""Evaluation players on the B200 engine.

* `get_ai_move(gs, model)` — scripts/play_vs_model.py:34-49: softmax(policy), restricted to the legal moves of the
  reference rules, argmax.  Same signature; `gs` is a `knightvision_b200.GameState` (or anything with the reference's
  `board` / `getValidMoves()` attributes), the forward pass runs on the tcgen05 network.
* `get_mcts_move(gs, model, sims)` — the same decision made by a PUCT search of `sims` simulations from that single
  position.  One game cannot fill a network batch with one simulation per wave, so the search runs `inflight`
  simulations per wave with virtual loss (kv_mcts_create_k).
* `arena(model_a, model_b, n_games, sims)` — the role of scripts/stockfish_play.py:26-111 (a strength estimate between
  training iterations) without an external engine or python-chess: two networks play each other under the reference's
  own rules, all games concurrently, each side searching with its own network.  Returns wins/draws/losses from A's side.

All computation is on the GPU (no CPU fallback).
"""
from __future__ import annotations

import numpy as np
import torch

from . import layout as L
from .ai import encode_board, encode_move
from .engine import Engine, lines_to_device
from .model import ChessNet
from .selfplay import engine_for


def _net_of(model) -> ChessNet:
    inner = model.module if hasattr(model, "module") else model
    if not isinstance(inner, ChessNet):
        net = ChessNet()
        net.load_state_dict(inner.state_dict())
        inner = net
    return inner.eval()


def get_ai_move(gs, model):
    valid_moves = gs.getValidMoves()
    if not valid_moves:
        return None
    net = _net_of(model)
    x = torch.from_numpy(np.asarray(encode_board(gs.board), dtype=np.float32)[None]).cuda()
    policy_logits, _ = net(x)
    policy = torch.softmax(policy_logits[0], dim=0)
    idx = torch.tensor([encode_move(m.startRow, m.startCol, m.endRow, m.endCol) for m in valid_moves], device=policy.device)
    probs = policy[idx]
    if float(probs.sum()) == 0.0:
        return valid_moves[0]                        # play_vs_model.py:44-46
    return valid_moves[int(torch.argmax(probs))]     # first maximum, as list.index(max(...)) does


def get_mcts_move(gs, model, sims: int = 800, inflight: int = 32, seed: int = 0, engine: Engine | None = None):
    """Move with the most visits after `sims` PUCT simulations from gs (no root noise, no temperature)."""
    valid_moves = gs.getValidMoves()
    if not valid_moves:
        return None
    eng = engine or engine_for(torch.device("cuda", torch.cuda.current_device()))
    net = _net_of(model)
    net.attach(eng, max_batch=max(inflight, 2))
    eng.mcts_create(1, sims, max_plies=1, temp_plies=0, dir_eps=0.0, seed=seed, eval_mode=1, inflight=inflight)
    eng.mcts_reset(lines_to_device(gs._line()[None], eng.device), game_id_base=0)
    eng.mcts_run_sims(1)
    while eng.mcts_status()["sims_in_move"] < sims:
        eng.mcts_run_sims(4)
    root = eng.mcts_read_root(0)
    if len(root["N"]) == 0:
        return valid_moves[0]
    best = int(root["moves"][int(np.argmax(root["N"]))])      # first maximum = lowest move index on ties
    for m in valid_moves:
        if m.word() == best:
            return m
    return valid_moves[0]


def arena(model_a, model_b, n_games: int = 256, sims: int = 100, max_plies: int = 200, seed: int = 7, device=None,
          random_start_plies: int = 8) -> dict:
    """n_games games between two networks, A playing white in the even-numbered games and black in the odd ones.
    Both searches run on one engine: before every ply the side to move's weights are committed (device to device) and
    the positions of the games where that network is to move are searched together."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    eng = engine_for(dev)
    nets = [_net_of(model_a).to(dev), _net_of(model_b).to(dev)]
    if nets[0].arch != nets[1].arch:
        raise ValueError("arena: both networks must have the same architecture")
    for net in nets:
        net.attach(eng, max_batch=max(n_games, 2))        # same geometry: the second attach only hands weights over
    lines = eng.random_positions(n_games, max(random_start_plies, 1), seed) if random_start_plies else \
        lines_to_device(np.stack([L.start_line()] * n_games), dev)
    a_is_white = (torch.arange(n_games, device=dev) % 2) == 0
    result = torch.zeros(n_games, dtype=torch.int32, device=dev)     # from white's side: +1 / -1 / 0
    done = torch.zeros(n_games, dtype=torch.bool, device=dev)
    eng.mcts_create(n_games, sims, max_plies=1, temp_plies=0, dir_eps=0.0, seed=seed, eval_mode=1)
    for ply in range(max_plies):
        if bool(done.all()):
            break
        white_to_move = (lines[:, 12] & 1) == 1
        for who in (0, 1):                                            # network A's boards, then network B's
            mine = (~done) & (white_to_move == a_is_white if who == 0 else white_to_move != a_is_white)
            if not bool(mine.any()):
                continue
            nets[who].sync_weights()                                  # device to device; also clears the evaluation cache
            # every game is searched; the move is kept only where this network is to move (simple, 2x the minimum work)
            eng.mcts_reset(lines.contiguous(), game_id_base=ply * n_games)
            eng.mcts_run_move()
            lines = torch.where(mine[:, None], eng.mcts_roots(), lines)
        moves, counts, flags = eng.movegen(lines)
        torch.cuda.synchronize()
        stm_white = (lines[:, 12] & 1) == 1
        mate = (counts == 0) & ((flags & 1) != 0)
        over = (counts == 0) | ((flags & 16) != 0)
        newly = over & ~done
        result = torch.where(newly & mate, torch.where(stm_white, -1, 1).to(torch.int32), result)
        done |= over
    res_a = torch.where(a_is_white, result, -result)
    return {"games": n_games, "a_wins": int((res_a > 0).sum()), "b_wins": int((res_a < 0).sum()),
            "draws": int((res_a == 0).sum()), "unfinished": int((~done).sum())}
