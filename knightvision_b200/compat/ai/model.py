"""ai.model of the reference tree -> the B200 network."""
from knightvision_b200.model import ChessNet  # noqa: F401
