"""Single-board Python surface of the reference rules engine (core/chessEngine.py: GameState :20, Move :683,
CastleRights :13) on top of the B200 kernels.

This shim exists so scripts written against the reference (`gs.getValidMoves()`, `gs.makeMove(m)`, attribute reads
and writes such as `gs.board[r][c] = 'wp'`) run unchanged; the batched device API (knightvision_b200.engine) is
the fast path.  All rules computation — legal move generation in reference order, make-move, squareUnderAttack —
happens in libkv_b200.so through its host-buffer entry points (kv_movegen_host / kv_make_moves_host); this file
only converts between the reference's attribute layout and the 128-byte board line and keeps the logs the
reference keeps (moveLog, enPassantPossibleLog, halfMoveClockLog, positionCounts).  No CUDA device ⇒ every rules
method raises.
"""
from __future__ import annotations

import numpy as np

from . import layout as L

_engine = None


def _eng():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(0)
    return _engine


def set_engine(engine):
    """Use an existing Engine (one kv_ctx per GPU) for the single-board calls."""
    global _engine
    _engine = engine


class CastleRights:
    def __init__(self, wks, wqs, bks, bqs):          # argument order of core/chessEngine.py:13-18
        self.wks, self.wqs, self.bks, self.bqs = wks, wqs, bks, bqs


class Move:
    ranksToRows = {"1": 7, "2": 6, "3": 5, "4": 4, "5": 3, "6": 2, "7": 1, "8": 0}
    rowsToRanks = {v: k for k, v in ranksToRows.items()}
    filesToCols = {"a": 0, "b": 1, "c": 2, "d": 3, "e": 4, "f": 5, "g": 6, "h": 7}
    colsToFiles = {v: k for k, v in filesToCols.items()}

    def __init__(self, startSq, endSq, board, isCastleMove=False, isEnPassantMove=False):
        self.startRow, self.startCol = startSq
        self.endRow, self.endCol = endSq
        self.pieceMoved = board[self.startRow][self.startCol]
        self.pieceCaptured = board[self.endRow][self.endCol]
        self.isEnPassantMove = isEnPassantMove
        self.enPassantPossible = ()
        if isEnPassantMove:
            self.pieceCaptured = "bp" if self.pieceMoved == "wp" else "wp"
        self.isPawnPromotion = (self.pieceMoved == "wp" and self.endRow == 0) or \
                               (self.pieceMoved == "bp" and self.endRow == 7)
        self.moveID = self.startRow * 1000 + self.startCol * 100 + self.endRow * 10 + self.endCol
        self.promotionChoice = "Q"
        self.isCastleMove = isCastleMove

    __hash__ = None   # the reference defines __eq__ without __hash__: Move objects are unhashable

    def __eq__(self, other):
        return (isinstance(other, Move) and self.startRow == other.startRow and self.startCol == other.startCol
                and self.endRow == other.endRow and self.endCol == other.endCol
                and self.pieceMoved == other.pieceMoved and self.isEnPassantMove == other.isEnPassantMove)

    def getChessNotation(self):
        return self.getRankFile(self.startRow, self.startCol) + self.getRankFile(self.endRow, self.endCol)

    def getRankFile(self, r, c):
        return self.colsToFiles[c] + self.rowsToRanks[r]

    def word(self) -> int:
        return L.move_word(self.startRow, self.startCol, self.endRow, self.endCol, self.isEnPassantMove,
                           bool(getattr(self, "isCastleMove", False)), self.isPawnPromotion)


_FLAG_ATTRS = (("wKingMoved", L.F_WK), ("bKingMoved", L.F_BK), ("wRookKingsideMoved", L.F_WRK),
               ("wRookQueensideMoved", L.F_WRQ), ("bRookKingsideMoved", L.F_BRK), ("bRookQueensideMoved", L.F_BRQ))


class GameState:
    def __init__(self):
        self.board = [row[:] for row in L.START_BOARD]
        self.whiteToMove = True
        self.moveLog = []
        self.whiteKingLocation = (7, 4)
        self.blackKingLocation = (0, 4)
        self.insideSquareUnderAttack = False
        self.checkMate = False
        self.staleMate = False
        for a, _ in _FLAG_ATTRS:
            setattr(self, a, False)
        self.enPassantPossible = ()
        self.enPassantPossibleLog = []
        self.halfMoveClockLog = []
        self.moveLogHistory = []
        self.boardHistory = {}
        self.halfMoveClock = 0
        self.boardStateCounter = {}
        self.draw50 = False
        self.drawRepetition = False
        self.positionCounts = {}

    # ---- attribute layout <-> board line -------------------------------------------------------------
    def _line(self) -> np.ndarray:
        moved = 0
        for a, bit in _FLAG_ATTRS:
            if getattr(self, a):
                moved |= bit
        return L.pack_fields(self.board, self.whiteToMove, self.whiteKingLocation, self.blackKingLocation, moved,
                             self.enPassantPossible, self.halfMoveClock, fen_pawns=True)

    def _adopt(self, line, board_only=False):
        f = L.unpack_fields(line)
        self.board = f["board"]
        if board_only:
            return
        self.whiteToMove = f["white_to_move"]
        self.whiteKingLocation, self.blackKingLocation = f["wk"], f["bk"]
        for a, bit in _FLAG_ATTRS:
            setattr(self, a, bool(f["moved"] & bit))
        self.enPassantPossible = f["ep"]
        self.halfMoveClock = f["clock"]

    # ---- rules (device) ------------------------------------------------------------------------------------
    def getValidMoves(self):
        moves, counts, flags, after = _eng().movegen_host(self._line()[None])
        fl = int(flags[0])
        if fl & L.RF_STATE_MUTATED:          # getKingMoves' restore quirk rewrote the board (:564)
            self._adopt(after[0], board_only=True)
        out = []
        for w in moves[0, :min(int(counts[0]), moves.shape[1])]:
            sr, sc, er, ec, ep, castle, _promo = L.move_fields(int(w))
            out.append(Move((sr, sc), (er, ec), self.board, isCastleMove=castle, isEnPassantMove=ep))
        self.checkMate = bool(fl & L.RF_CHECKMATE)
        self.staleMate = bool(fl & L.RF_STALEMATE)
        self.draw50 = bool(fl & L.RF_DRAW50)
        # checkForEndConditions :632-651: repetition is looked at only when none of the above fired
        self.drawRepetition = bool(out) and not self.draw50 and self.positionCounts.get(self.getFEN(), 0) >= 3
        return out

    def makeMove(self, move):
        self.enPassantPossibleLog.append(self.enPassantPossible)
        self.halfMoveClockLog.append(self.halfMoveClock)
        self.enPassantPossibleLog.append(self.enPassantPossible)     # the reference logs it twice (:129, :167)
        key = self.getBoardStateKey
        new = _eng().make_moves_host(self._line()[None], np.array([move.word()], dtype=np.uint16))[0]
        self._adopt(new)
        if move.isPawnPromotion and move.promotionChoice != "Q":     # :190-191 honours a caller-set choice
            self.board[move.endRow][move.endCol] = move.pieceMoved[0] + move.promotionChoice
        self.moveLog.append(move)
        k = str(self.board) + str(not self.whiteToMove)              # getBoardStateKey before the side flip (:183)
        self.boardStateCounter[k] = self.boardStateCounter.get(k, 0) + 1
        fen = self.getFEN()
        self.positionCounts[fen] = self.positionCounts.get(fen, 0) + 1
        del key

    def undoMove(self):
        """Inverse of makeMove with the reference's observable behaviour (:202-271): board, side, king location,
        e.p. square and clock come back; the moved-flag of a king/corner-rook move is cleared unconditionally;
        positionCounts is never decremented."""
        if not self.moveLog:
            return
        self.enPassantPossible = self.enPassantPossibleLog.pop() if self.enPassantPossibleLog else ()
        if self.halfMoveClockLog:
            self.halfMoveClock = self.halfMoveClockLog.pop()
        m = self.moveLog.pop()
        b = self.board
        b[m.startRow][m.startCol] = m.pieceMoved
        b[m.endRow][m.endCol] = m.pieceCaptured
        if m.isEnPassantMove:
            b[m.endRow][m.endCol] = "--"
            b[m.startRow][m.endCol] = m.pieceCaptured
        if m.pieceMoved == "wK":
            self.wKingMoved = False
            self.whiteKingLocation = (m.startRow, m.startCol)
        elif m.pieceMoved == "bK":
            self.bKingMoved = False
            self.blackKingLocation = (m.startRow, m.startCol)
        elif m.pieceMoved == "wR":
            if (m.startRow, m.startCol) == (7, 0):
                self.wRookQueensideMoved = False
            elif (m.startRow, m.startCol) == (7, 7):
                self.wRookKingsideMoved = False
        elif m.pieceMoved == "bR":
            if (m.startRow, m.startCol) == (0, 0):
                self.bRookQueensideMoved = False
            elif (m.startRow, m.startCol) == (0, 7):
                self.bRookKingsideMoved = False
        if getattr(m, "isCastleMove", False):
            if m.endCol - m.startCol == 2:
                b[m.endRow][m.endCol + 1] = b[m.endRow][m.endCol - 1]
                b[m.endRow][m.endCol - 1] = "--"
            else:
                b[m.endRow][m.endCol - 2] = b[m.endRow][m.endCol + 1]
                b[m.endRow][m.endCol + 1] = "--"
        self.whiteToMove = not self.whiteToMove
        self.enPassantPossible = self.enPassantPossibleLog.pop() if self.enPassantPossibleLog else ()

    def squareUnderAttack(self, r, c):
        return bool(_eng().attacked_host(self._line()[None])[0] >> (r * 8 + c) & 1)

    def inCheck(self):
        k = self.whiteKingLocation if self.whiteToMove else self.blackKingLocation
        return self.squareUnderAttack(k[0], k[1])

    # ---- host bookkeeping (no rules computation) ---------------------------------------------------------------
    def isDraw(self):
        """:21-33 — the 50-move clause is unsatisfiable (pieceMoved is never '--'); only-kings is what is left."""
        pieces = {p for row in self.board for p in row if p != "--"}
        return pieces <= {"wK", "bK"}

    def getFEN(self):
        rows = []
        for row in self.board:
            s, empty = "", 0
            for sq in row:
                if sq == "--":
                    empty += 1
                    continue
                if empty:
                    s += str(empty)
                    empty = 0
                s += sq[1].upper() if sq[0] == "w" else sq[1].lower()
            if empty:
                s += str(empty)
            rows.append(s)
        return "/".join(rows) + (" w" if self.whiteToMove else " b")

    def getBoardStateKey(self):
        return str(self.board) + str(self.whiteToMove)

    def loadFEN(self, fen):
        """:85-122 — sets board, side, e.p. and an unused castleRights; NOT the king locations, moved-flags or clock.
        Like the reference it writes pawns as 'wP' / 'bP' (`char.upper()`, :100), so `board` and getFEN() agree with
        it; the rules kernels treat those as ordinary pawns (the reference treats them as a separate kind that moves
        like a pawn but never promotes, captures e.p. or gives a pawn check — an accident no reference caller relies
        on: nothing calls loadFEN).  Documented in INTEGRATION.md."""
        parts = fen.split()
        for r, txt in enumerate(parts[0].split("/")[:8]):
            row = []
            for ch in txt:
                if ch.isdigit():
                    row.extend(["--"] * int(ch))
                else:
                    row.append(("w" if ch.isupper() else "b") + ch.upper())
            self.board[r] = row
        self.whiteToMove = parts[1] == "w"
        if not hasattr(self, "castleRights"):
            self.castleRights = CastleRights(False, False, False, False)
        cr = parts[2]
        self.castleRights.wks, self.castleRights.bks = "K" in cr, "k" in cr
        self.castleRights.wqs, self.castleRights.bqs = "Q" in cr, "q" in cr
        if parts[3] != "-":
            self.enPassantPossible = (8 - int(parts[3][1]), ord(parts[3][0]) - ord("a"))
        else:
            self.enPassantPossible = ()
        self.moveLog = []
        self.enPassantPossibleLog = []
