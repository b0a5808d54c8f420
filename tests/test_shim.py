"""The reference's own unit tests (tests/test_castling.py, test_en_passant.py, test_promotion.py) restated against
the knightvision_b200 GameState / Move shim, which computes through the C ABI on the GPU."""
import json
import os

import numpy as np
import pytest

import helpers as H
from knightvision_b200 import layout as L
from oracle import kv_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def engine():
    from knightvision_b200 import chess_engine as CE
    from knightvision_b200.engine import Engine
    e = Engine(0)
    CE.set_engine(e)
    yield e
    e.close()


def _castle_state(white):
    from knightvision_b200 import GameState
    gs = GameState()
    gs.board = [["--"] * 8 for _ in range(8)]
    gs.board[0][0] = gs.board[0][7] = "bR"; gs.board[0][4] = "bK"
    gs.board[7][0] = gs.board[7][7] = "wR"; gs.board[7][4] = "wK"
    gs.whiteKingLocation, gs.blackKingLocation = (7, 4), (0, 4)
    for a in ("wKingMoved", "bKingMoved", "wRookKingsideMoved", "wRookQueensideMoved", "bRookKingsideMoved",
              "bRookQueensideMoved"):
        setattr(gs, a, False)
    gs.whiteToMove = white
    return gs


@pytest.mark.parametrize("white,row", [(True, 7), (False, 0)])
def test_castling_both_wings(white, row):          # reference tests/test_castling.py:32-56
    from knightvision_b200 import Move
    gs = _castle_state(white)
    moves = gs.getValidMoves()
    assert Move((row, 4), (row, 6), gs.board, isCastleMove=True) in moves
    assert Move((row, 4), (row, 2), gs.board, isCastleMove=True) in moves
    gold = json.load(open(os.path.join(H.GOLDEN, "unit_tests.json")))["castle_w" if white else "castle_b"]
    assert [m.word() for m in moves] == gold["moves"]          # full list, reference order


def test_en_passant_make_and_undo():                # reference tests/test_en_passant.py:17-43
    from knightvision_b200 import GameState, Move
    gs = GameState()
    gs.board = [["--"] * 8 for _ in range(8)]
    gs.board[7][4] = "wK"; gs.board[0][4] = "bK"; gs.board[3][4] = "wp"; gs.board[1][3] = "bp"
    gs.whiteToMove = False
    gs.makeMove(Move((1, 3), (3, 3), gs.board))
    assert gs.enPassantPossible == (2, 3)
    ep = Move((3, 4), (2, 3), gs.board, isEnPassantMove=True)
    assert ep in gs.getValidMoves()
    gs.makeMove(ep)
    assert gs.board[2][3] == "wp" and gs.board[3][3] == "--" and gs.board[3][4] == "--"
    gs.undoMove()
    assert gs.board[3][4] == "wp" and gs.board[3][3] == "bp" and gs.board[2][3] == "--"
    assert gs.enPassantPossible == (2, 3)
    gs.undoMove()
    assert gs.board[1][3] == "bp" and gs.board[3][3] == "--" and gs.enPassantPossible == ()


@pytest.mark.parametrize("white", [True, False])
def test_promotion_make_and_undo(white):            # reference tests/test_promotion.py:9-39
    from knightvision_b200 import GameState, Move
    gs = GameState()
    if white:
        gs.board[6][0] = "--"; gs.board[1][0] = "wp"
        mv = Move((1, 0), (0, 0), gs.board)
        mv.isPawnPromotion = True; mv.promotionChoice = "Q"
        gs.makeMove(mv)
        assert gs.board[0][0] == "wQ"
        gs.undoMove()
        assert gs.board[1][0] == "wp" and gs.board[0][0] == "bR"
    else:
        gs.board[1][7] = "--"; gs.board[6][7] = "bp"; gs.whiteToMove = False
        mv = Move((6, 7), (7, 7), gs.board)
        mv.isPawnPromotion = True; mv.promotionChoice = "N"
        gs.makeMove(mv)
        assert gs.board[7][7] == "bN"
        gs.undoMove()
        assert gs.board[6][7] == "bp" and gs.board[7][7] == "wR"


def test_shim_playout_matches_oracle_and_flags():
    from knightvision_b200 import GameState
    rng = np.random.default_rng(2)
    gs = GameState()
    line = L.start_line()
    for ply in range(60):
        moves = gs.getValidMoves()
        em, ec, ef, mid = O.movegen(line[None].copy())
        assert [m.word() for m in moves] == [int(x) for x in em[0, :ec[0]]]
        assert gs.checkMate == bool(ef[0] & L.RF_CHECKMATE) and gs.staleMate == bool(ef[0] & L.RF_STALEMATE)
        assert gs.inCheck() == O.in_check(mid[0])
        if not moves:
            break
        m = moves[int(rng.integers(len(moves)))]
        gs.makeMove(m)
        line = O.make_moves(mid, np.array([m.word()], dtype=np.uint16))[0]
        assert np.array_equal(gs._line()[:13], line[:13])
        assert gs.getFEN().split()[1] == ("w" if gs.whiteToMove else "b")
    assert len(gs.moveLog) == ply + (1 if moves else 0)


def test_encoders_match_reference_goldens():
    from knightvision_b200 import decode_move_index, encode_board, encode_move
    g = np.load(H.GOLDEN + "/encode.npz")
    for i in (0, 17, 63):
        board = L.unpack_fields(g["lines"][i])["board"]
        assert np.array_equal(encode_board(board), g["planes"][i])
    assert encode_move(6, 0, 5, 0) == 3112 and encode_move(6, 4, 4, 4) == 3364 and encode_move(7, 6, 5, 5) == 4013
    for i, sr, sc, er, ec in g["decode"]:
        assert decode_move_index(int(i)) == (sr, sc, er, ec)

