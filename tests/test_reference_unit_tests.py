"""The reference's own eight unit tests, run VERBATIM (tests/golden/ref_tests/ holds its three test files byte for byte)
against the B200 shim through the `chessEngine` alias module, and the E9 bookkeeping (positionCounts, drawRepetition,
getFEN, loadFEN, undoMove) against sequences recorded from the unmodified reference (tests/golden/gamestate_seq.json,
oracle/gen_golden_selfplay.py; core/chessEngine.py:85-122, :193-197, :202-271, :632-678).  Rules calls go through the
C ABI on the GPU."""
import sys

import pytest

import helpers_shim as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def engine():
    from knightvision_b200 import chess_engine as CE
    from knightvision_b200 import compat
    from knightvision_b200.engine import Engine
    e = Engine(0)
    CE.set_engine(e)
    sys.path.insert(0, compat.PATH)            # what the reference's tests do with its core/ directory
    yield e
    sys.path.remove(compat.PATH)
    S.purge_alias_modules()
    e.close()


def test_reference_unit_tests_verbatim():
    S.run_reference_unit_tests()


def test_position_counts_repetition_fen_sequences():
    S.run_gamestate_sequences()


def test_load_fen_matches_reference():
    S.run_load_fen()


def test_reference_import_paths_resolve_to_the_engine():
    """scripts/self_play.py:19,87-92 / scripts/learn.py:34 import lines, with knightvision_b200/compat first on sys.path."""
    S.purge_alias_modules()
    from ai import encode_board, encode_move          # noqa: F401
    from ai.model import ChessNet
    from core.chessEngine import GameState
    from scripts.self_play import generate_self_play_data, self_play   # noqa: F401
    import knightvision_b200 as KV
    assert GameState is KV.GameState and ChessNet is KV.ChessNet and self_play is KV.self_play
    assert encode_move(6, 4, 4, 4) == 3364
    S.purge_alias_modules()
