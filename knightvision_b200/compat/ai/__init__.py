"""ai package of the reference tree (ai/__init__.py:1 re-exports the encoders) -> the B200 implementations."""
from knightvision_b200.ai import decode_move_index, encode_board, encode_move  # noqa: F401
