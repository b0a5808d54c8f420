"""Host-side description of the 128-byte board line that the CUDA kernels operate on.

One board = 16 little-endian u64 words (exactly one 128 B cache line; lane i of the warp that owns
the board loads word i):

  w[0..11]  piece bitboards in the reference's PIECE_TO_INDEX order (ai/ai.py:7-10):
            wK wQ wR wB wN wp bK bQ bR bB bN bp; bit index = row*8 + col, row 0 = rank 8
            (the reference's board[row][col] indexing, core/chessEngine.py:39-47)
  w[12]     meta: bit 0 whiteToMove | bits 1-6 moved flags (wK,bK,wRk,wRq,bRk,bRq; :66-71)
            | bits 8-14 enPassantPossible square (64 = none; :72) | bits 16-21 whiteKingLocation
            | bits 24-29 blackKingLocation (:59-60, independent of the board) | bits 32-47 halfMoveClock (:79)
  w[13..15] reserved, zero

This module is pure host logic (numpy); it performs no rules computation.
"""
from __future__ import annotations

import numpy as np

PIECES = ("wK", "wQ", "wR", "wB", "wN", "wp", "bK", "bQ", "bR", "bB", "bN", "bp")
PIECE_TO_INDEX = {p: i for i, p in enumerate(PIECES)}
LINE_WORDS = 16
EP_NONE = 64

F_WK, F_BK, F_WRK, F_WRQ, F_BRK, F_BRQ = 1, 2, 4, 8, 16, 32

# move word (u16): from | to<<6 | ep<<12 | castle<<13 | promo<<14
MF_EP, MF_CASTLE, MF_PROMO = 1, 2, 4
# result flags returned by move generation
RF_CHECKMATE, RF_STALEMATE, RF_DRAW50, RF_E3_CHECK, RF_ONLY_KINGS, RF_STATE_MUTATED, RF_OVERFLOW = (
    1, 2, 4, 8, 16, 32, 64)

START_BOARD = [
    ["bR", "bN", "bB", "bQ", "bK", "bB", "bN", "bR"],
    ["bp"] * 8,
    ["--"] * 8, ["--"] * 8, ["--"] * 8, ["--"] * 8,
    ["wp"] * 8,
    ["wR", "wN", "wB", "wQ", "wK", "wB", "wN", "wR"],
]


# core/chessEngine.py:85-122 loadFEN writes pawns as 'wP' / 'bP' (upper-case kind): a 13th / 14th piece kind that moves
# like a pawn (:49) but is no pawn for promotion, e.p. capture, pawn checks or the encoder.  The line has no code for it.
FEN_PAWNS = {"wP": "wp", "bP": "bp"}


def pack_fields(board, white_to_move=True, wk=(7, 4), bk=(0, 4), moved=0, ep=(), clock=0, fen_pawns=False) -> np.ndarray:
    """Pack reference-style fields into one line (np.uint64[16]).  Unknown piece codes are ignored (what the reference's
    encoder does, ai/ai.py:26-29); with fen_pawns the 'wP' / 'bP' codes loadFEN produces count as ordinary pawns."""
    w = np.zeros(LINE_WORDS, dtype=np.uint64)
    bbs = [0] * 12
    for r in range(8):
        row = board[r]
        for c in range(8):
            code = row[c]
            if fen_pawns:
                code = FEN_PAWNS.get(code, code)
            idx = PIECE_TO_INDEX.get(code)
            if idx is not None:
                bbs[idx] |= 1 << (r * 8 + c)
    for i in range(12):
        w[i] = np.uint64(bbs[i])
    ep_sq = EP_NONE if not ep else ep[0] * 8 + ep[1]
    meta = (int(bool(white_to_move)) | ((moved & 63) << 1) | (ep_sq << 8) | ((wk[0] * 8 + wk[1]) << 16)
            | ((bk[0] * 8 + bk[1]) << 24) | (min(int(clock), 0xFFFF) << 32))
    w[12] = np.uint64(meta)
    return w


def unpack_fields(w) -> dict:
    """Inverse of pack_fields."""
    w = [int(x) for x in np.asarray(w, dtype=np.uint64)]
    board = [["--"] * 8 for _ in range(8)]
    for i, name in enumerate(PIECES):
        bb = w[i]
        while bb:
            low = bb & -bb
            sq = low.bit_length() - 1
            board[sq >> 3][sq & 7] = name
            bb ^= low
    m = w[12]
    ep_sq = (m >> 8) & 127
    wk, bk = (m >> 16) & 63, (m >> 24) & 63
    return dict(board=board, white_to_move=bool(m & 1), moved=(m >> 1) & 63,
                ep=() if ep_sq >= 64 else (ep_sq >> 3, ep_sq & 7),
                wk=(wk >> 3, wk & 7), bk=(bk >> 3, bk & 7), clock=(m >> 32) & 0xFFFF)


def start_line() -> np.ndarray:
    return pack_fields(START_BOARD)


def move_word(sr, sc, er, ec, ep=False, castle=False, promo=False) -> int:
    return (sr * 8 + sc) | ((er * 8 + ec) << 6) | (int(ep) << 12) | (int(castle) << 13) | (int(promo) << 14)


def move_fields(mv: int):
    f, t = mv & 63, (mv >> 6) & 63
    return (f >> 3, f & 7, t >> 3, t & 7, bool(mv >> 12 & 1), bool(mv >> 13 & 1), bool(mv >> 14 & 1))


def move_uci(mv: int) -> str:
    sr, sc, er, ec = move_fields(mv)[:4]
    return "abcdefgh"[sc] + str(8 - sr) + "abcdefgh"[ec] + str(8 - er)
