#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) in the build container.

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together
with this script.  Run:  python oracle/gen_golden.py [--skip-d5] [--skip-net]

Outputs (tests/golden/):
  perft.json        perft counts / category counts / per-root-move divide + FNV order digests for the
                    start position (d1-5) and the six positions of the reference's own tests (d1-4)
  playouts.npz      positions visited by seeded random playouts of the reference engine: packed line
                    before/after getValidMoves, legal move words in reference order, result flags,
                    inCheck(), squareUnderAttack() over all 64 squares, the move played, line after makeMove
  synthetic.npz     same, for seeded synthetic ("sane" and "wild") boards that exercise the quirk list
  encode.npz        encode_board / encode_move goldens (ai/ai.py)
  net.npz           ChessNet(seed 0) fp32 CPU outputs on 8 playout positions (ai/model.py)
  unit_tests.json   the reference's 8 unit tests restated as data (tests/test_*.py)
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REF, "core"))

from knightvision_b200 import layout as L  # noqa: E402

import chessEngine as CE  # noqa: E402  (the unmodified reference engine, as its own tests import it)

FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
M64 = (1 << 64) - 1
GOLD = os.path.join(REPO, "tests", "golden")


# ------------------------------------------------------------------------------------------------
def gs_to_line(gs) -> np.ndarray:
    moved = ((L.F_WK if gs.wKingMoved else 0) | (L.F_BK if gs.bKingMoved else 0)
             | (L.F_WRK if gs.wRookKingsideMoved else 0) | (L.F_WRQ if gs.wRookQueensideMoved else 0)
             | (L.F_BRK if gs.bRookKingsideMoved else 0) | (L.F_BRQ if gs.bRookQueensideMoved else 0))
    return L.pack_fields(gs.board, gs.whiteToMove, gs.whiteKingLocation, gs.blackKingLocation, moved,
                         gs.enPassantPossible, gs.halfMoveClock)


def line_to_gs(line):
    f = L.unpack_fields(line)
    gs = CE.GameState()
    gs.board = f["board"]
    gs.whiteToMove = f["white_to_move"]
    gs.whiteKingLocation, gs.blackKingLocation = f["wk"], f["bk"]
    m = f["moved"]
    gs.wKingMoved, gs.bKingMoved = bool(m & L.F_WK), bool(m & L.F_BK)
    gs.wRookKingsideMoved, gs.wRookQueensideMoved = bool(m & L.F_WRK), bool(m & L.F_WRQ)
    gs.bRookKingsideMoved, gs.bRookQueensideMoved = bool(m & L.F_BRK), bool(m & L.F_BRQ)
    gs.enPassantPossible = f["ep"]
    gs.halfMoveClock = f["clock"]
    return gs


def mv_word(m) -> int:
    return L.move_word(m.startRow, m.startCol, m.endRow, m.endCol, m.isEnPassantMove, m.isCastleMove,
                       m.isPawnPromotion)


def fnv_moves(h, moves):
    for m in moves:
        for b in (m.startRow * 8 + m.startCol, m.endRow * 8 + m.endCol,
                  int(m.isEnPassantMove) | (int(m.isCastleMove) << 1) | (int(m.isPawnPromotion) << 2)):
            h = ((h ^ b) * FNV_PRIME) & M64
    return h


def clone(gs):
    """Field-wise clone (never relies on undoMove)."""
    n = CE.GameState()
    n.board = [row[:] for row in gs.board]
    for a in ("whiteToMove", "whiteKingLocation", "blackKingLocation", "wKingMoved", "bKingMoved",
              "wRookKingsideMoved", "wRookQueensideMoved", "bRookKingsideMoved", "bRookQueensideMoved",
              "enPassantPossible", "halfMoveClock"):
        setattr(n, a, getattr(gs, a))
    return n


def perft(gs, depth, acc):
    """acc = [nodes, cap, ep, castle, promo, digest]; copy-make DFS, bulk count at depth 1."""
    moves = gs.getValidMoves()
    acc[5] = fnv_moves(acc[5], moves)
    if depth == 1:
        acc[0] += len(moves)
        for m in moves:
            acc[1] += m.pieceCaptured != "--"
            acc[2] += bool(m.isEnPassantMove)
            acc[3] += bool(m.isCastleMove)
            acc[4] += bool(m.isPawnPromotion)
        return
    for m in moves:
        ch = clone(gs)
        ch.makeMove(m)
        perft(ch, depth - 1, acc)


def _perft_child(args):
    line, mv, depth = args
    gs = line_to_gs(line)
    for m in gs.getValidMoves():
        if mv_word(m) == mv:
            gs.makeMove(m)
            break
    else:
        raise RuntimeError("move not found")
    acc = [0, 0, 0, 0, 0, FNV_OFFSET]
    perft(gs, depth, acc)
    return [int(x) for x in acc]


def test_positions():
    """The start position and the six positions the reference's tests construct."""
    out = []
    out.append(("startpos", CE.GameState(), 5))
    # tests/test_castling.py:9-30
    def castle(white):
        gs = CE.GameState()
        gs.board = [["--"] * 8 for _ in range(8)]
        gs.board[0][0] = gs.board[0][7] = "bR"; gs.board[0][4] = "bK"
        gs.board[7][0] = gs.board[7][7] = "wR"; gs.board[7][4] = "wK"
        gs.whiteToMove = white
        return gs
    out.append(("castle_w", castle(True), 4))
    out.append(("castle_b", castle(False), 4))
    # tests/test_en_passant.py:5-13,:18-21,:46-48
    def ep(a):
        gs = CE.GameState()
        gs.board = [["--"] * 8 for _ in range(8)]
        gs.board[7][4] = "wK"; gs.board[0][4] = "bK"
        if a:
            gs.board[3][4] = "wp"; gs.board[1][3] = "bp"; gs.whiteToMove = False
        else:
            gs.board[6][3] = "wp"; gs.board[4][4] = "bp"
        return gs
    out.append(("ep_a", ep(True), 4))
    out.append(("ep_b", ep(False), 4))
    # tests/test_promotion.py:10-14,:26-30
    gs = CE.GameState(); gs.board[6][0] = "--"; gs.board[1][0] = "wp"
    out.append(("promo_w", gs, 4))
    gs = CE.GameState(); gs.board[1][7] = "--"; gs.board[6][7] = "bp"; gs.whiteToMove = False
    out.append(("promo_b", gs, 4))
    return out


def gen_perft(skip_d5):
    res = {}
    pool = mp.Pool(os.cpu_count())
    for name, gs, maxd in test_positions():
        line = gs_to_line(gs)
        root_moves = gs.getValidMoves()
        root_words = [mv_word(m) for m in root_moves]
        entry = {"line": [int(x) for x in line], "root_moves": root_words,
                 "root_uci": [m.getChessNotation() for m in root_moves],
                 "root_digest": fnv_moves(FNV_OFFSET, root_moves), "depths": {}}
        acc = [0, 0, 0, 0, 0, FNV_OFFSET]
        perft(clone(gs), 1, acc)
        entry["depths"]["1"] = {"nodes": acc[0], "cats": acc[1:5]}
        for d in range(2, maxd + 1):
            if d == 5 and skip_d5:
                continue
            parts = pool.map(_perft_child, [(line, w, d - 1) for w in root_words])
            entry["depths"][str(d)] = {
                "nodes": sum(p[0] for p in parts),
                "cats": [sum(p[k] for p in parts) for k in range(1, 5)],
                "divide": [p[0] for p in parts],
                "child_digests": [p[5] for p in parts],
            }
            print(name, d, entry["depths"][str(d)]["nodes"], flush=True)
        res[name] = entry
    pool.close()
    with open(os.path.join(GOLD, "perft.json"), "w") as f:
        json.dump(res, f)


# ------------------------------------------------------------------------------------------------
def observe(gs):
    """Run the reference on a position; returns a dict of everything we pin."""
    line_in = gs_to_line(gs)
    sua = 0
    for r in range(8):
        for c in range(8):
            if gs.squareUnderAttack(r, c):
                sua |= 1 << (r * 8 + c)
    incheck = gs.inCheck()
    e3 = gs.checkForPinsAndChecks()[0]
    moves = gs.getValidMoves()
    line_mid = gs_to_line(gs)
    flags = ((L.RF_CHECKMATE if gs.checkMate else 0) | (L.RF_STALEMATE if gs.staleMate else 0)
             | (L.RF_DRAW50 if gs.draw50 else 0) | (L.RF_E3_CHECK if e3 else 0))
    pieces = {p for row in gs.board for p in row if p != "--"}
    if pieces <= {"wK", "bK"}:
        flags |= L.RF_ONLY_KINGS
    if not np.array_equal(line_in[:12], line_mid[:12]):
        flags |= L.RF_STATE_MUTATED
    return dict(line_in=line_in, line_mid=line_mid, sua=sua, incheck=int(incheck), flags=flags,
                moves=moves, words=[mv_word(m) for m in moves])


def pick_move(rng, gs, moves, mode):
    if mode == "aggressive":
        kingcaps = [m for m in moves if m.pieceCaptured[1:] == "K"]
        if kingcaps:
            return rng.choice(kingcaps)
        # chase the unseen knight check square (king + (-2,+1)), SURVEY Q1
        ek = gs.blackKingLocation if gs.whiteToMove else gs.whiteKingLocation
        tgt = (ek[0] - 2, ek[1] + 1)
        kn = [m for m in moves if m.pieceMoved[1] == "N" and (m.endRow, m.endCol) == tgt]
        if kn and rng.random() < 0.9:
            return rng.choice(kn)
        caps = [m for m in moves if m.pieceCaptured != "--"]
        if caps and rng.random() < 0.5:
            return rng.choice(caps)
    if mode == "pawnrush":
        pw = [m for m in moves if m.pieceMoved[1] == "p"]
        if pw and rng.random() < 0.6:
            return rng.choice(pw)
    return rng.choice(moves)


def gen_playouts(n_games, seed, max_plies):
    rng = random.Random(seed)
    rows = []
    for g in range(n_games):
        mode = ("uniform", "aggressive", "pawnrush")[g % 3]
        gs = CE.GameState()
        for ply in range(max_plies):
            ob = observe(gs)
            moves = ob["moves"]
            if not moves:
                ob["played"] = 0xFFFF
                ob["line_out"] = ob["line_mid"]
                rows.append(ob)
                break
            m = pick_move(rng, gs, moves, mode)
            ob["played"] = mv_word(m)
            gs.makeMove(m)
            ob["line_out"] = gs_to_line(gs)
            rows.append(ob)
            if gs.isDraw():
                break
    return rows


def save_rows(rows, path):
    n = len(rows)
    maxm = max(len(r["words"]) for r in rows)
    moves = np.zeros((n, maxm), dtype=np.uint16)
    for i, r in enumerate(rows):
        moves[i, :len(r["words"])] = r["words"]
    np.savez_compressed(
        path,
        line_in=np.stack([r["line_in"] for r in rows]),
        line_mid=np.stack([r["line_mid"] for r in rows]),
        line_out=np.stack([r["line_out"] for r in rows]),
        sua=np.array([r["sua"] for r in rows], dtype=np.uint64),
        incheck=np.array([r["incheck"] for r in rows], dtype=np.uint8),
        flags=np.array([r["flags"] for r in rows], dtype=np.int32),
        counts=np.array([len(r["words"]) for r in rows], dtype=np.int32),
        moves=moves,
        played=np.array([r["played"] for r in rows], dtype=np.uint16),
    )
    print(path, n, "rows; max moves", maxm, flush=True)


def synthetic_state(rng, wild):
    gs = CE.GameState()
    b = [["--"] * 8 for _ in range(8)]
    squares = [(r, c) for r in range(8) for c in range(8)]
    rng.shuffle(squares)
    it = iter(squares)
    wk = next(it); bk = next(it)
    have_wk = have_bk = True
    if wild and rng.random() < 0.25:
        have_wk = False
    if wild and rng.random() < 0.25:
        have_bk = False
    if have_wk:
        b[wk[0]][wk[1]] = "wK"
    if have_bk:
        b[bk[0]][bk[1]] = "bK"
    if not wild and rng.random() < 0.5:
        # bias towards home-square kings and corner rooks so castling paths are exercised
        b = [["--"] * 8 for _ in range(8)]
        wk, bk = (7, 4), (0, 4)
        b[7][4] = "wK"; b[0][4] = "bK"
        for (r, c, p) in ((7, 0, "wR"), (7, 7, "wR"), (0, 0, "bR"), (0, 7, "bR")):
            if rng.random() < 0.8:
                b[r][c] = p
        squares = [(r, c) for r in range(8) for c in range(8) if b[r][c] == "--"]
        rng.shuffle(squares)
        it = iter(squares)
    n_extra = rng.randint(0, 14 if not wild else 20)
    for _ in range(n_extra):
        try:
            r, c = next(it)
        except StopIteration:
            break
        col = rng.choice("wb")
        t = rng.choice("QRBNppp")
        if t == "p" and not wild and r in (0, 7):
            continue
        b[r][c] = col + t
    gs.board = b
    gs.whiteToMove = rng.random() < 0.5
    gs.whiteKingLocation, gs.blackKingLocation = wk, bk
    if wild and rng.random() < 0.3:
        gs.whiteKingLocation = (rng.randrange(8), rng.randrange(8))
    if wild and rng.random() < 0.3:
        gs.blackKingLocation = (rng.randrange(8), rng.randrange(8))
    for a in ("wKingMoved", "bKingMoved", "wRookKingsideMoved", "wRookQueensideMoved",
              "bRookKingsideMoved", "bRookQueensideMoved"):
        setattr(gs, a, rng.random() < 0.25)
    if rng.random() < 0.4:
        if wild:
            gs.enPassantPossible = (rng.randrange(8), rng.randrange(8))
        else:
            # consistent e.p.: a pawn of the side that just moved stands in front of an empty square
            r = 2 if gs.whiteToMove else 5
            pr = 3 if gs.whiteToMove else 4
            c = rng.randrange(8)
            if b[r][c] == "--":
                b[pr][c] = "bp" if gs.whiteToMove else "wp"
                gs.enPassantPossible = (r, c)
    gs.halfMoveClock = rng.choice([0, 0, 3, 99, 100, 150])
    return gs


def gen_synthetic(n, seed):
    rng = random.Random(seed)
    rows = []
    for i in range(n):
        gs = synthetic_state(rng, wild=(i % 2 == 1))
        ob = observe(gs)
        if ob["moves"]:
            m = rng.choice(ob["moves"])
            ob["played"] = mv_word(m)
            gs.makeMove(m)
            ob["line_out"] = gs_to_line(gs)
        else:
            ob["played"] = 0xFFFF
            ob["line_out"] = ob["line_mid"]
        rows.append(ob)
    return rows


# ------------------------------------------------------------------------------------------------
def import_ai():
    """ai/* needs `chess`; core/__init__ needs pygame: stub both (SURVEY §8c route b)."""
    chess = types.ModuleType("chess")
    chess.SQUARES = range(64)
    chess.Board = type("Board", (), {})
    chess.WHITE, chess.PAWN = True, 1
    sys.modules.setdefault("chess", chess)
    sys.modules.setdefault("pygame", types.ModuleType("pygame"))
    sys.path.insert(0, REF)
    import ai.ai as refai
    from ai.model import ChessNet
    return refai, ChessNet


def gen_encode_and_net(rows, skip_net):
    refai, ChessNet = import_ai()
    rng = random.Random(7)
    idx = sorted(rng.sample(range(len(rows)), 64))
    lines = np.stack([rows[i]["line_in"] for i in idx])
    planes = np.stack([refai.encode_board(L.unpack_fields(l)["board"]) for l in lines])
    mv = []
    for i in idx:
        for w in rows[i]["words"][:4]:
            sr, sc, er, ec = L.move_fields(w)[:4]
            mv.append((w, refai.encode_move(sr, sc, er, ec)))
    dec = [(i,) + tuple(refai.decode_move_index(i)) for i in (0, 1, 63, 64, 512, 3112, 3364, 3902, 4013, 4095)]
    np.savez_compressed(os.path.join(GOLD, "encode.npz"), lines=lines, planes=planes.astype(np.float32),
                        move_words=np.array([m[0] for m in mv], dtype=np.uint16),
                        move_index=np.array([m[1] for m in mv], dtype=np.int32),
                        decode=np.array(dec, dtype=np.int32))
    if skip_net:
        return
    import torch
    torch.manual_seed(0)
    net = ChessNet().eval()
    # non-trivial BN running stats so folding is actually exercised
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.05)
            m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    x = torch.from_numpy(planes[:8])
    with torch.no_grad():
        pol, val = net(x)
    torch.manual_seed(0)
    net0 = ChessNet().eval()
    with torch.no_grad():
        pol0, val0 = net0(x)
    np.savez_compressed(os.path.join(GOLD, "net.npz"), lines=lines[:8],
                        policy_bnrand=pol.numpy(), value_bnrand=val.numpy(),
                        policy_init=pol0.numpy(), value_init=val0.numpy(),
                        n_params=np.array(sum(p.numel() for p in net.parameters())),
                        keys=np.array(list(net.state_dict().keys())))
    print("net goldens written", flush=True)


def gen_unit_tests():
    """tests/test_castling.py, test_en_passant.py, test_promotion.py restated as data."""
    out = {}
    for name, gs, _ in test_positions():
        if name.startswith("castle"):
            mv = gs.getValidMoves()
            out[name] = {"line": [int(x) for x in gs_to_line(gs)], "moves": [mv_word(m) for m in mv]}
    with open(os.path.join(GOLD, "unit_tests.json"), "w") as f:
        json.dump(out, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-d5", action="store_true")
    ap.add_argument("--skip-net", action="store_true")
    ap.add_argument("--skip-perft", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    rows = gen_playouts(n_games=45, seed=20261018, max_plies=160)
    save_rows(rows, os.path.join(GOLD, "playouts.npz"))
    syn = gen_synthetic(3000, seed=99)
    save_rows(syn, os.path.join(GOLD, "synthetic.npz"))
    gen_unit_tests()
    gen_encode_and_net(rows, args.skip_net)
    if not args.skip_perft:
        gen_perft(args.skip_d5)


if __name__ == "__main__":
    main()
