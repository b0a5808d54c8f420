"""knightvision_b200 — B200-native (sm_100a) self-play hot path for KnightVision.

Reference-compatible surface: GameState / Move (core/chessEngine.py), encode_board / encode_move /
decode_move_index (ai/ai.py), ChessNet (ai/model.py), self_play / generate_self_play_data (scripts/self_play.py).
Batched device API: knightvision_b200.engine.Engine.  Everything computes in libkv_b200.so (hand-written CUDA);
there is no CPU fallback.
"""
from . import layout  # noqa: F401

__all__ = ["GameState", "Move", "CastleRights", "ChessNet", "encode_board", "encode_move", "decode_move_index",
           "self_play", "generate_self_play_data", "Engine", "get_ai_move", "get_mcts_move", "arena"]


def __getattr__(name):
    # lazy: importing the package must not need torch / the .so (the CPU build check imports it)
    if name in ("GameState", "Move", "CastleRights"):
        from . import chess_engine
        return getattr(chess_engine, name)
    if name == "ChessNet":
        from .model import ChessNet
        return ChessNet
    if name in ("encode_board", "encode_move", "decode_move_index"):
        from . import ai
        return getattr(ai, name)
    if name in ("self_play", "generate_self_play_data", "SelfPlay"):
        from . import selfplay
        return getattr(selfplay, name)
    if name in ("get_ai_move", "get_mcts_move", "arena"):
        from . import players
        return getattr(players, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
