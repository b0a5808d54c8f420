"""`import chessEngine` alias (the reference's tests and GUI import core/chessEngine.py this way): put this directory
on sys.path instead of the reference's `core/` and GameState / Move / CastleRights are the B200 shim."""
from knightvision_b200.chess_engine import CastleRights, GameState, Move  # noqa: F401
