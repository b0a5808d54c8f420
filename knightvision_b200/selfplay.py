"""Drop-in for the reference's self-play entry points (scripts/self_play.py:258-311), on the B200 engine.

    self_play(model, num_games, device, max_moves=None, model_path=None) -> list[(np.float32[12,8,8], int, float)]
    generate_self_play_data(model, num_games, device, max_moves=None)     -> same, decisive-only filter (:304-310)

Every game runs concurrently on the GPU: bitboard rules kernels, a GPU-resident PUCT tree per game and the
tcgen05 policy/value tower evaluate one leaf per game per wave.  The reference picks each move by sampling the raw
policy; here the move comes from `KV_SIMS` PUCT simulations (default 800; the tree search is new functionality,
DESIGN.md §MCTS) with the reference's Dirichlet parameters (DIR_NOISE_EPS / DIR_NOISE_ALPHA, :12-13) at the root;
`KV_SIMS=1` is the reference's own rule (no search: the move is sampled from the noisy policy mixed as :150-167 does).
The game loop's stopping rules are the reference's, in its order (:180-199): only kings, resignation (value < -0.7 after
more than 15 moves), max_moves, then checkmate / stalemate.
Record format, reward map (win 1.0 / draw 0.2 / loss -1.0 from white's side, same on every ply, :245-250), error
behaviour (ValueError without model and path, FileNotFoundError for a missing checkpoint) follow the reference.
"""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from .engine import Engine
from .model import ChessNet

logger = logging.getLogger(__name__)

DIR_NOISE_EPS = float(os.getenv("DIR_NOISE_EPS", "0.25"))      # scripts/self_play.py:12
DIR_NOISE_ALPHA = float(os.getenv("DIR_NOISE_ALPHA", "0.3"))   # :13
SEED = int(os.getenv("SEED", "42"))                            # :24
DEFAULT_SIMS = int(os.getenv("KV_SIMS", "800"))
DEFAULT_PLY_CAP = int(os.getenv("KV_MAX_PLIES", "512"))        # the reference has no cap when max_moves is None
TEMP_PLIES = int(os.getenv("KV_TEMP_PLIES", "30"))
C_PUCT = float(os.getenv("KV_CPUCT", "1.5"))
RESIGN_THRESHOLD = float(os.getenv("KV_RESIGN_THRESHOLD", "-0.7"))   # scripts/self_play.py:185 (value < threshold ...
RESIGN_MIN_MOVES = int(os.getenv("KV_RESIGN_MIN_MOVES", "15"))       # ... and move_count > 15); a negative count = never

_engines: dict[int, Engine] = {}


def engine_for(device) -> Engine:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("knightvision_b200 self-play needs a CUDA device (there is no CPU fallback)")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]


def _load_model(model_path: str) -> ChessNet:
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"Model checkpoint not found: {model_path}")   # scripts/self_play.py:67-68
    ckpt = torch.load(model_path, map_location="cpu")
    sd = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt   # :72-76
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}               # DataParallel
    net = ChessNet()
    net.load_state_dict(sd)
    return net.eval()


class SelfPlay:
    """n_games concurrent self-play games on one GPU."""

    def __init__(self, model: ChessNet, n_games: int, device, sims: int = DEFAULT_SIMS, max_plies: int = DEFAULT_PLY_CAP,
                 temp_plies: int = TEMP_PLIES, c_puct: float = C_PUCT, dir_alpha: float = DIR_NOISE_ALPHA,
                 dir_eps: float = DIR_NOISE_EPS, seed: int = SEED, eval_mode: int = 1, engine: Engine | None = None,
                 inflight: int | None = None, resign_threshold: float = RESIGN_THRESHOLD,
                 resign_min_moves: int = RESIGN_MIN_MOVES):
        self.eng = engine or engine_for(device)
        self.n_games, self.sims, self.max_plies = n_games, sims, max_plies
        if inflight is None:
            # few games cannot fill a network batch with one simulation per game and wave (the reference's default is 5
            # games per generation, scripts/learn.py:106): run K simulations per game and wave with virtual loss so that
            # a wave carries ~2 048 leaves; 2 048 games or more keep the sequential search (K = 1)
            inflight = int(os.getenv("KV_INFLIGHT", "0")) or max(1, min(32, 2048 // max(n_games, 1), max(sims // 8, 1)))
        self.inflight = inflight
        if eval_mode == 1:
            inner = model.module if hasattr(model, "module") else model
            if not isinstance(inner, ChessNet):
                # a reference ai.model.ChessNet (or any module with its state_dict layout): adopt its weights
                net = ChessNet()
                net.load_state_dict(inner.state_dict())
                inner = net
            inner.eval()
            inner.attach(self.eng, max_batch=max(n_games * inflight, 2))
            self.model = inner
        geom = (n_games, sims, max_plies, temp_plies, c_puct, dir_alpha, dir_eps, seed, eval_mode, inflight)
        if getattr(self.eng, "mcts_geometry_key", None) != geom:      # same search context: keep the pools (and the cache)
            self.eng.mcts_create(n_games, sims, max_plies, temp_plies, c_puct, dir_alpha, dir_eps, seed, eval_mode,
                                 inflight=inflight)
            self.eng.mcts_geometry_key = geom
        self.eng.mcts_set_resign(resign_threshold, resign_min_moves)

    def play(self, start_lines: torch.Tensor | None = None, game_id_base: int = 0, progress=None) -> dict:
        eng = self.eng
        eng.mcts_reset(start_lines, game_id_base)
        st = eng.mcts_status()
        moves = 0
        while st["done"] < self.n_games and moves < self.max_plies + 1:
            eng.mcts_run_move()
            moves += 1
            st = eng.mcts_status()
            if progress:
                progress(moves, st)
        return st

    def records_device(self):
        return self.eng.mcts_records()

    def records(self):
        """The reference's list of (state float32 (12,8,8), move_index, reward) tuples, in game order."""
        lines, move, reward, game = self.eng.mcts_records()
        if lines.shape[0] == 0:
            return []
        return records_to_tuples(self.eng, lines, move, reward)


def records_to_tuples(eng, lines, move, reward):
    """Packed device records -> the reference's list of (np.float32 (12,8,8), int, float) tuples (scripts/self_play.py:253):
    planes by the encode kernel, one device-to-host copy per field, the tuples zipped from views of the host arrays."""
    planes = eng.encode(lines.contiguous()).cpu().numpy()
    return list(zip(planes, move.cpu().tolist(), reward.cpu().tolist()))


def self_play(model, num_games, device, max_moves=None, model_path=None):
    if model is None and model_path is None:
        raise ValueError("Either model or model_path must be provided")   # scripts/self_play.py:259-260
    if model is None:
        model = _load_model(model_path)
    sp = SelfPlay(model, int(num_games), device, max_plies=int(max_moves) if max_moves else DEFAULT_PLY_CAP)
    st = sp.play()
    logger.info("self-play: %d games, %d plies, W/B/D = %d/%d/%d", num_games, st["plies"], st["white_wins"],
                st["black_wins"], st["draws"])
    return sp.records()


MIN_DECISIVE_GAMES = 10   # scripts/self_play.py:305 (it counts records, not games)


def filter_decisive(data):
    """scripts/self_play.py:303-310: keep only the records of won / lost games (reward +-1.0) once there are at least
    MIN_DECISIVE_GAMES of them; otherwise return everything."""
    decisive = [s for s in data if s[2] == 1.0 or s[2] == -1.0]
    return decisive if len(decisive) >= MIN_DECISIVE_GAMES else data


def filter_decisive_device(lines, move, reward, game=None):
    """The same filter on the packed device records (learn loop: no host round trip)."""
    mask = reward.abs() == 1.0
    if int(mask.sum()) < MIN_DECISIVE_GAMES:
        return (lines, move, reward) if game is None else (lines, move, reward, game)
    out = (lines[mask], move[mask], reward[mask])
    return out if game is None else out + (game[mask],)


def generate_self_play_data(model, num_games, device, max_moves=None):
    return filter_decisive(self_play(model, num_games, device, max_moves))
