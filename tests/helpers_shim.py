"""Shared bodies of the shim tests: run on the B200 through the C ABI (tests/test_reference_unit_tests.py, -m gpu) and,
for the host-side bookkeeping alone, on the CPU with the rules calls answered by the pinned oracle
(tests/test_shim_host_logic.py)."""
import importlib.util
import json
import os
import sys
import unittest

import numpy as np

import helpers as H
from knightvision_b200 import layout as L

REF_TESTS = os.path.join(H.GOLDEN, "ref_tests")


class OracleEngine:
    """Stand-in for knightvision_b200.engine.Engine in CPU tests of the shim's HOST logic: the three rules calls the shim
    makes are answered by the oracle (test infrastructure; the product never does this)."""

    def movegen_host(self, lines, stride=256):
        from oracle import kv_oracle as O
        m, c, f, after = O.movegen(np.array(lines, dtype=np.uint64, copy=True))
        return m, c, f, after

    def make_moves_host(self, lines, mv):
        from oracle import kv_oracle as O
        return O.make_moves(lines, mv)

    def attacked_host(self, lines):
        return H.oracle_attack_masks(np.ascontiguousarray(lines, dtype=np.uint64))


def run_reference_unit_tests():
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    for name in ("test_castling", "test_en_passant", "test_promotion"):
        spec = importlib.util.spec_from_file_location("kv_ref_" + name, os.path.join(REF_TESTS, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite.addTests(loader.loadTestsFromModule(mod))
    assert suite.countTestCases() == 8
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    assert res.wasSuccessful() and res.testsRun == 8, (res.failures, res.errors)
    import chessEngine
    from knightvision_b200 import chess_engine as CE
    assert chessEngine.GameState is CE.GameState and chessEngine.Move is CE.Move


def _check_step(gs, st, where):
    assert gs.getFEN() == st["fen"], where
    fen = st["fen"]
    assert gs.positionCounts.get(fen, 0) == st["count_cur"], where
    assert len(gs.positionCounts) == st["count_keys"] and sum(gs.positionCounts.values()) == st["count_sum"], where
    assert gs.halfMoveClock == st["clock"] and gs.isDraw() == st["is_draw"], where
    assert [int(x) for x in gs._line()[:13]] == st["line"][:13], where
    n = len(gs.getValidMoves())
    assert n == st["n_moves"], where
    assert (gs.checkMate, gs.staleMate, gs.draw50, gs.drawRepetition) == \
        (st["checkMate"], st["staleMate"], st["draw50"], st["drawRepetition"]), where
    assert gs.inCheck() == st["in_check"], where


def run_gamestate_sequences():
    from knightvision_b200 import GameState
    gold = json.load(open(os.path.join(H.GOLDEN, "gamestate_seq.json")))
    saw_rep = saw_undo = 0
    for seq in gold["sequences"]:
        gs = GameState()
        for i, st in enumerate(seq["steps"]):
            where = (seq["name"], i, st["op"])
            if st["op"] == "move":
                mv = [m for m in gs.getValidMoves() if m.word() == st["word"]]
                assert len(mv) == 1 and mv[0].getChessNotation() == st["uci"], where
                gs.makeMove(mv[0])
            elif st["op"] == "undo":
                gs.undoMove()
                saw_undo += 1
            _check_step(gs, st, where)
            saw_rep += int(st["drawRepetition"])
        assert dict(gs.positionCounts) == seq["final_counts"], seq["name"]
    assert saw_rep > 0 and saw_undo > 0


def run_load_fen():
    from knightvision_b200 import GameState
    gold = json.load(open(os.path.join(H.GOLDEN, "gamestate_seq.json")))
    for st in gold["load_fen"]:
        gs = GameState()
        gs.loadFEN(st["fen_in"])
        assert gs.getFEN() == st["fen"] and list(gs.enPassantPossible) == st["ep"]
        assert list(gs.whiteKingLocation) == st["wk"] and list(gs.blackKingLocation) == st["bk"]   # NOT set by loadFEN (Q14)
        assert gs.board == st["board"]                        # incl. the 'wP' / 'bP' codes of :100
        if any(p in ("wP", "bP") for row in st["board"] for p in row):
            # documented deviation: the kernels play FEN pawns as ordinary pawns (the reference's 'P' kind never promotes,
            # captures e.p. or checks); same moves as the board with lower-case pawns
            twin = GameState()
            twin.board = [[L.FEN_PAWNS.get(p, p) for p in row] for row in gs.board]
            twin.whiteToMove, twin.enPassantPossible = gs.whiteToMove, gs.enPassantPossible
            assert [m.word() for m in gs.getValidMoves()] == [m.word() for m in twin.getValidMoves()]
        else:
            assert [m.word() for m in gs.getValidMoves()] == st["moves"]
            assert [int(x) for x in gs._line()[:13]] == st["line"][:13]


def purge_alias_modules():
    for k in list(sys.modules):
        if k == "chessEngine" or k.split(".")[0] in ("core", "ai", "scripts"):
            sys.modules.pop(k)
