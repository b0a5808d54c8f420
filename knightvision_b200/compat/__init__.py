"""Import-path compatibility with the reference tree.  With this directory first on sys.path,

    from chessEngine import GameState, Move            (reference tests/, core/chessMain.py)
    from core.chessEngine import GameState             (scripts/self_play.py:87)
    from ai import encode_board, encode_move           (scripts/self_play.py:19,88)
    from ai.model import ChessNet                      (scripts/self_play.py:92)
    from scripts.self_play import self_play, generate_self_play_data   (scripts/learn.py:34, scripts/train.py:37)

resolve to the B200 implementations, so the reference's callers run unchanged."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
