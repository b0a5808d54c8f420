"""core.chessEngine of the reference tree -> the B200 shim."""
from knightvision_b200.chess_engine import CastleRights, GameState, Move  # noqa: F401
