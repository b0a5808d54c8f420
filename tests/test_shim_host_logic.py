"""Host-side logic of the GameState / Move shim on the CPU: the reference's eight unit tests verbatim and the E9
bookkeeping sequences (positionCounts, drawRepetition, getFEN, loadFEN, undoMove logs), with the shim's three rules calls
answered by the pinned oracle instead of the GPU (tests/helpers_shim.OracleEngine).  The same bodies run through the
C ABI on the B200 in tests/test_reference_unit_tests.py."""
import sys

import pytest

import helpers_shim as S


@pytest.fixture(scope="module", autouse=True)
def engine():
    from knightvision_b200 import chess_engine as CE
    from knightvision_b200 import compat
    saved = CE._engine
    CE.set_engine(S.OracleEngine())
    sys.path.insert(0, compat.PATH)
    yield
    sys.path.remove(compat.PATH)
    S.purge_alias_modules()
    CE.set_engine(saved)


def test_reference_unit_tests_verbatim_host_logic():
    S.run_reference_unit_tests()


def test_position_counts_repetition_fen_sequences_host_logic():
    S.run_gamestate_sequences()


def test_load_fen_host_logic():
    S.run_load_fen()


def test_castle_rights_argument_order():
    from knightvision_b200 import CastleRights
    cr = CastleRights(True, False, False, True)        # (wks, wqs, bks, bqs), core/chessEngine.py:13-18
    assert (cr.wks, cr.wqs, cr.bks, cr.bqs) == (True, False, False, True)
