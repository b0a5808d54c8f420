"""Board / move encoders with the reference's names and semantics (ai/ai.py:17-57)."""
from __future__ import annotations

import numpy as np

from . import layout as L

PIECE_TO_INDEX = dict(L.PIECE_TO_INDEX)                      # ai/ai.py:7-10
INDEX_TO_PIECE = {v: k for k, v in PIECE_TO_INDEX.items()}


def encode_board(board) -> np.ndarray:
    """One-hot float32 [12,8,8] planes of a reference-style board (list of lists / ndarray of 2-char codes).
    Runs the device encoder (kv_encode); unknown codes are ignored like the reference's dict lookup."""
    from .chess_engine import _eng
    from .engine import lines_to_device
    eng = _eng()
    rows = [list(r) for r in (board.tolist() if isinstance(board, np.ndarray) else board)]
    line = L.pack_fields(rows)[None]
    return eng.encode(lines_to_device(line, eng.device))[0].cpu().numpy()


def decode_move_index(index):
    start, end = index // 64, index % 64
    return (start // 8, start % 8, end // 8, end % 8)


def encode_move(start_row, start_col, end_row, end_col):
    return (start_row * 8 + start_col) * 64 + (end_row * 8 + end_col)


__all__ = ["encode_board", "decode_move_index", "encode_move"]
