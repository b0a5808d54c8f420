// kv_rules.cuh — KnightVision's custom chess rules as bitboard device code, W lanes per board (W = 16 or 32).
//
// A warp holds 32 / W boards; the W lanes of a group own one board.  The 128-byte board line (16 x u64, see
// include/kv_b200.h) is loaded with one coalesced request per board (lane q of the group loads word q) and broadcast
// inside the group by shuffles; every lane then holds the whole position in registers.  Work is spread over the
// group's lanes by *work item*:
//   - checkForPinsAndChecks  (core/chessEngine.py:325-383): lanes 0-7 walk one king ray each
//   - getKingMoves/getCastleMoves (:543-601): lanes 0-7 test one king step each on the "king placed"
//     board, lanes 8-12 test the castle squares, all through attacked()
//   - getAllPossibleMoves (:433-441): lane q owns the q-th piece of the side to move in square order (the row-major
//     scan order of the reference), so every lane that works has a piece; more than W pieces take another round.
//     The move list falls out of a group prefix sum over the per-piece counts, each piece emitting in the reference's
//     direction order
// The rules kernels (perft, movegen, make-move) run W = 16 (two boards per warp: half the warp instructions per
// board); the tree-search kernels, which own one game per warp, use W = 32 through the *_warp wrappers.
// squareUnderAttack (:400-415) — "some opponent pseudo-move ENDS on the square", with all its quirks
// (pawn pushes attack, pawn diagonals onto empty squares do not, opponent castling attacks c/g,
// nested calls see the re-entrancy guard) — is restated in closed form in attacked().
// The quirk list this file reproduces on purpose is SURVEY.md §8a-Q.
//
// Every warp collective sits in code that all 32 lanes reach together (loops whose trip count depends on a board run
// until no board of the warp needs another pass).
// The same source compiles under KV_HOST_EMU (tests/simt_emu) where the warp collectives are routed
// to a 32-fiber lock-step emulator, so the integer kernels are parity-checked on a box without a GPU.
#pragma once
#include "kv_tables.cuh"
#include "kv_warp.cuh"

namespace kv {

constexpr int F_WK = 1, F_BK = 2, F_WRK = 4, F_WRQ = 8, F_BRK = 16, F_BRQ = 32;   // core/chessEngine.py:66-71
constexpr int MF_EP = 1, MF_CASTLE = 2, MF_PROMO = 4;
constexpr int RF_CHECKMATE = 1, RF_STALEMATE = 2, RF_DRAW50 = 4, RF_E3_CHECK = 8, RF_ONLY_KINGS = 16,
              RF_STATE_MUTATED = 32, RF_OVERFLOW = 64;
constexpr int T_K = 0, T_Q = 1, T_R = 2, T_B = 3, T_N = 4, T_P = 5;               // ai/ai.py:7-10 order
constexpr int MAX_MOVES = 256;
constexpr int LINE_WORDS = 16;
constexpr int EP_NONE = 64;

constexpr uint64_t FILE_A = 0x0101010101010101ull;
constexpr uint64_t FILE_H = FILE_A << 7;

KV_DEV uint64_t bit(int s) { return 1ull << s; }
KV_DEV bool dir_asc(int d) { return (0xCA >> d) & 1; }   // S, E, SW, SE walk towards higher square indices
KV_DEV int dir_opp(int d) { return d < 4 ? (d ^ 1) : 11 - d; }
KV_DEV int first_on_ray(uint64_t b, int d) { return dir_asc(d) ? ctz64(b) : msb64(b); }

// Squares a slider on s reaches in direction d (up to and including the first blocker).
// Branch-free: towards higher squares the nearest blocker is the lowest set bit l of (ray & occ) and the reach is the
// ray up to l (l ^ (l - 1) = all bits up to l; all ones when there is no blocker); towards lower squares it is the
// highest set bit (bit 0 stands in when there is none: the mask is then all ones as well).
KV_DEV uint64_t ray_att(const Tables& T, int d, int s, uint64_t occ) {
    const uint64_t r = T.ray[d][s];
    const uint64_t b = r & occ;
    if (dir_asc(d)) {
        const uint64_t l = b & (0ull - b);
        return r & (l ^ (l - 1ull));
    }
    return r & (~0ull << msb64(b | 1ull));
}
KV_DEV uint64_t rook_att(const Tables& T, int s, uint64_t occ) {
    return ray_att(T, 0, s, occ) | ray_att(T, 1, s, occ) | ray_att(T, 2, s, occ) | ray_att(T, 3, s, occ);
}
KV_DEV uint64_t bishop_att(const Tables& T, int s, uint64_t occ) {
    return ray_att(T, 4, s, occ) | ray_att(T, 5, s, occ) | ray_att(T, 6, s, occ) | ray_att(T, 7, s, occ);
}

// What attacked() needs to know about a (possibly hypothetical) board: "own" = side to move.
struct Agg {
    uint64_t occ, own, opp;
    uint64_t eN, eK, eRQ, eBQ, eP, eR;
};

// squareUnderAttack(r,c), core/chessEngine.py:400-415, for attacker = opponent of the side to move:
// true iff getAllPossibleMoves() of the opponent (pins=[], nested guard on) contains a move ending on t.
//   wtm    side to move is white (so the attacker is black and its pawns move towards higher rows)
//   ep     enPassantPossible square or EP_NONE;  moved: moved-flag bits;  akloc: attacker's king *location variable*
KV_DEV bool attacked(const Tables& T, const Agg& g, int t, bool wtm, int ep, int moved, int akloc) {
    const uint64_t tb = bit(t);
    bool res = false;
    if (!(g.opp & tb)) {   // knight / king / slider moves end on empty or enemy squares only
        uint64_t a = (T.knight[t] & g.eN) | (T.king[t] & g.eK);
        a |= rook_att(T, t, g.occ) & g.eRQ;
        a |= bishop_att(T, t, g.occ) & g.eBQ;
        res = a != 0;
    }
    uint64_t capsrc, p1src, p2src, mid;
    bool row2;
    if (wtm) {   // black pawns: moveAmount +1, startRow 1 (core/chessEngine.py:452-455)
        capsrc = ((tb >> 7) & ~FILE_A) | ((tb >> 9) & ~FILE_H);
        p1src = tb >> 8;
        p2src = tb >> 16;
        mid = tb >> 8;
        row2 = (t >> 3) == 3;
    } else {     // white pawns: moveAmount -1, startRow 6
        capsrc = ((tb << 7) & ~FILE_H) | ((tb << 9) & ~FILE_A);
        p1src = tb << 8;
        p2src = tb << 16;
        mid = tb << 8;
        row2 = (t >> 3) == 4;
    }
    // pawn capture: target holds a piece of the side to move, or is the e.p. square (:466-472)
    if ((capsrc & g.eP) && ((g.own & tb) || t == ep)) res = true;
    // pawn pushes count as "attacks" on empty squares (:459-462)
    if (!(g.occ & tb)) {
        if (p1src & g.eP) res = true;
        if (row2 && (p2src & g.eP) && !(g.occ & mid)) res = true;
    }
    // the opponent's available castling "attacks" its king-destination square (:575-601);
    // getKingMoves is only reached when the board scan finds an opponent king
    if (g.eK) {
        if (wtm) {
            if (akloc == 4 && !(moved & F_BK)) {
                if (t == 6 && !(moved & F_BRK) && !(g.occ & (bit(5) | bit(6))) && (g.eR & bit(7))) res = true;
                if (t == 2 && !(moved & F_BRQ) && !(g.occ & (bit(1) | bit(2) | bit(3))) && (g.eR & bit(0))) res = true;
            }
        } else {
            if (akloc == 60 && !(moved & F_WK)) {
                if (t == 62 && !(moved & F_WRK) && !(g.occ & (bit(61) | bit(62))) && (g.eR & bit(63))) res = true;
                if (t == 58 && !(moved & F_WRQ) && !(g.occ & (bit(57) | bit(58) | bit(59))) && (g.eR & bit(56))) res = true;
            }
        }
    }
    return res;
}

// A position as every lane of the owning warp sees it.
struct Pos {
    uint64_t o[6];   // side to move: K Q R B N P
    uint64_t e[6];   // opponent
    uint64_t meta;
    bool wtm;
    int ep, moved, kloc, akloc, clock;
};

KV_DEV void make_agg(const Pos& p, Agg& g) {
    g.own = p.o[0] | p.o[1] | p.o[2] | p.o[3] | p.o[4] | p.o[5];
    g.opp = p.e[0] | p.e[1] | p.e[2] | p.e[3] | p.e[4] | p.e[5];
    g.occ = g.own | g.opp;
    g.eN = p.e[T_N];
    g.eK = p.e[T_K];
    g.eRQ = p.e[T_R] | p.e[T_Q];
    g.eBQ = p.e[T_B] | p.e[T_Q];
    g.eP = p.e[T_P];
    g.eR = p.e[T_R];
}

// The board line as the W lanes of a group hold it: lane q keeps words q, q + WL, ... (WL = min(W, 16) words per "row"):
// one word per lane for W >= 16 (lanes 16-31 of a W = 32 group hold nothing), two for W = 8 (words q and q + 8).
template <int W>
struct Line {
    static constexpr int WL = W < 16 ? W : 16;
    static constexpr int NW = 16 / WL;
    uint64_t w[NW];
};
// word i of the line, broadcast to every lane of the group (i is a compile-time constant at every call site)
template <int W>
KV_DEV uint64_t line_word(const Line<W>& L, int i, int lane) {
    return sub_shfl64<W>(L.w[i / Line<W>::WL], i % Line<W>::WL, lane);
}

// Broadcast the line to every lane of the group, side to move first.
template <int W>
KV_DEV void load_pos(const Line<W>& L, Pos& p, int lane) {
    p.meta = line_word<W>(L, 12, lane);
    p.wtm = p.meta & 1;
    if (Line<W>::NW == 1) {
        // the shuffle source lane is chosen by the side bit, so no selects afterwards
        const int so = p.wtm ? 0 : 6, se = 6 - so;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            p.o[i] = sub_shfl64<W>(L.w[0], so + i, lane);
            p.e[i] = sub_shfl64<W>(L.w[0], se + i, lane);
        }
    } else {
        // W = 8: bitboards 0-7 are word 0 of lanes 0-7, bitboards 8-11 word 1 of lanes 0-3.  White's six boards are
        // 0-5 (word 0), black's 6-11 (word 0 of lanes 6, 7; word 1 of lanes 0-3): each lane offers the word the side
        // bit asks for, the source lane is computed from it
        const uint64_t w0 = L.w[0], w1 = L.w[Line<W>::NW - 1];
        const uint64_t vo = p.wtm ? w0 : w1, ve = p.wtm ? w1 : w0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            p.o[i] = sub_shfl64<W>(w0, p.wtm ? i : 6 + i, lane);
            p.e[i] = sub_shfl64<W>(w0, p.wtm ? 6 + i : i, lane);
        }
#pragma unroll
        for (int i = 2; i < 6; i++) {
            p.o[i] = sub_shfl64<W>(vo, p.wtm ? i : i - 2, lane);
            p.e[i] = sub_shfl64<W>(ve, p.wtm ? i - 2 : i, lane);
        }
    }
    p.moved = (int)((p.meta >> 1) & 63);
    p.ep = (int)((p.meta >> 8) & 127);
    if (p.ep > EP_NONE) p.ep = EP_NONE;
    int wk = (int)((p.meta >> 16) & 63), bk = (int)((p.meta >> 24) & 63);
    p.kloc = p.wtm ? wk : bk;
    p.akloc = p.wtm ? bk : wk;
    p.clock = (int)((p.meta >> 32) & 0xFFFF);
}

struct KingOut {
    uint32_t steps;    // bit k: king step k (getKingMoves order) is legal
    uint32_t castle;   // bit0 king side, bit1 queen side
};

// getKingMoves + getCastleMoves for the king-like piece on ks (core/chessEngine.py:543-601), plus the
// extra in-check filter of getValidMoves (:306-309) when mode == 1.  Warp-collective; result is uniform.
template <int W>
KV_DEV KingOut king_phase(const Tables& T, int lane_, const Pos& p, const Agg& g, int ks, int mode,
                          uint64_t valid, uint64_t pinned) {
    const int home = p.wtm ? 60 : 4;
    const int f_k = p.wtm ? F_WK : F_BK, f_rk = p.wtm ? F_WRK : F_BRK, f_rq = p.wtm ? F_WRQ : F_BRQ;
    // castling can only come out if the king stands at home unmoved and a wing has its flag clear, its squares empty
    // and its rook in the corner; only then are the five squares of work items 8-12 worth testing
    const bool wing_k = !(p.moved & f_rk) && !(g.occ & (bit(home + 1) | bit(home + 2))) && (p.o[T_R] & bit(home + 3));
    const bool wing_q = !(p.moved & f_rq) && !(g.occ & (bit(home - 1) | bit(home - 2) | bit(home - 3))) &&
                        (p.o[T_R] & bit(home - 4));
    const bool castle_possible = p.kloc == home && !(p.moved & f_k) && (wing_k || wing_q);
    // 13 work items, W per round: items 0-7 test a king step on the "king placed" board (:556-563), item 8 the king's
    // own square, items 9,10 f,g and 11,12 c,d on the unmodified board (:576-599).  One convergent attacked() per round.
    constexpr int ROUNDS = (13 + W - 1) / W;
    uint32_t okb = 0, ab = 0;
#pragma unroll
    for (int rr = 0; rr < ROUNDS; rr++) {
        const int lane = sub_q<W>(lane_) + rr * W;   // work item index
        if (ROUNDS > 1 && rr > 0 && ballot(castle_possible) == 0) break;   // no board of the warp can castle
        Agg h = g;
        int t = 0;
        bool active = false;
        if (lane < 8) {
            // step k of getKingMoves' order (:544-546): row offset + 1 = {0,0,0,1,1,2,2,2}, column offset + 1 = {0,1,2,0,2,0,1,2}
            const int dr = (int)((0xA940u >> (2 * lane)) & 3u) - 1;
            const int dc = (int)((0x9224u >> (2 * lane)) & 3u) - 1;
            const int er = (ks >> 3) + dr, ec = (ks & 7) + dc;
            if (er >= 0 && er < 8 && ec >= 0 && ec < 8) {
                t = er * 8 + ec;
                const uint64_t tb = bit(t), ksb = bit(ks);
                if (!(g.own & tb)) {
                    active = true;   // the king placed on t, whatever stood there removed
                    h.occ = (g.occ & ~ksb) | tb;
                    h.own = (g.own & ~ksb) | tb;
                    h.opp = g.opp & ~tb;
                    h.eN = g.eN & ~tb;
                    h.eK = g.eK & ~tb;
                    h.eRQ = g.eRQ & ~tb;
                    h.eBQ = g.eBQ & ~tb;
                    h.eP = g.eP & ~tb;
                    h.eR = g.eR & ~tb;
                }
            }
        } else if (lane < 13) {
            active = castle_possible;
            t = lane == 8 ? ks : (lane == 9 ? home + 1 : (lane == 10 ? home + 2 : (lane == 11 ? home - 2 : home - 1)));
        }
        bool a1 = false;
        if (active) a1 = attacked(T, h, t, p.wtm, p.ep, p.moved, p.akloc);
        bool ok = lane < 8 && active && !a1;
        const bool att = lane >= 8 && lane < 13 && a1;
        if (mode == 1) {   // getValidMoves :306-309 re-tests king moves on the unmodified board
            bool a2 = false;
            if (ok) a2 = attacked(T, g, t, p.wtm, p.ep, p.moved, p.akloc);
            ok = ok && !a2;
        }
        // item index -> bit index: steps in bits 0-7 of okb, items 8-12 in bits 0-4 of ab
        const uint32_t bo = sub_ballot<W>(ok, lane_), ba = sub_ballot<W>(att, lane_);
        if (W >= 16) {
            okb = bo & 0xFFu;
            ab = ba >> 8;
        } else if (rr == 0) {
            okb = bo & 0xFFu;
        } else {
            ab = ba & 0x1Fu;
        }
    }
    KingOut out;
    out.steps = okb;
    out.castle = 0;
    if (castle_possible && !(ab & 1)) {
        bool ck = wing_k && !(ab & 2) && !(ab & 4);
        bool cq = wing_q && !(ab & 8) && !(ab & 16);
        if (mode == 1) {
            // getValidMoves :306-310: Move.pieceMoved is board[home]; a king (either colour) is re-tested
            // with squareUnderAttack(to) — already false here — anything else must land on validSquares
            const bool home_is_k = ((p.o[T_K] | p.e[T_K]) & bit(home)) != 0;
            if (!home_is_k) {
                ck = ck && (valid & bit(home + 2));
                cq = cq && (valid & bit(home - 2));
            }
        }
        out.castle = (ck ? 1u : 0u) | (cq ? 2u : 0u);
    }
    if (pinned & bit(ks)) {
        // a king piece standing on a ray from a stale king location can be "pinned" (:604-630): colinearity filter
        int pd = 0;
        for (int d = 0; d < 8; d++)
            if (T.ray[d][p.kloc] & bit(ks)) pd = d;
        const uint64_t line = T.ray[pd][ks] | T.ray[dir_opp(pd)][ks];
        uint32_t keep = 0;
        for (int k = 0; k < 8; k++) {
            const int dr = (k < 3) ? -1 : (k < 5 ? 0 : 1);
            const int dc = (k == 0 || k == 3 || k == 5) ? -1 : ((k == 1 || k == 6) ? 0 : 1);
            const int er = (ks >> 3) + dr, ec = (ks & 7) + dc;
            if (er >= 0 && er < 8 && ec >= 0 && ec < 8 && (line & bit(er * 8 + ec))) keep |= 1u << k;
        }
        out.steps &= keep;
        if (!(line & bit(home + 2))) out.castle &= ~1u;
        if (!(line & bit(home - 2))) out.castle &= ~2u;
    }
    return out;
}

// Destination squares of the legal king steps (bit k of `steps` = step k of getKingMoves' order; only on-board steps are
// ever set).  The eight steps form a 3 x 3 block around ks: lay them out as three board rows with the king's square at
// bit 9 and shift the block to ks.
KV_DEV uint64_t king_steps_to_mask(int ks, uint32_t steps) {
    const uint32_t pat = (steps & 7u) | ((steps & 8u) << 5) | ((steps & 16u) << 6) | ((steps & 0xE0u) << 11);
    return ks >= 9 ? ((uint64_t)pat << (ks - 9)) : ((uint64_t)pat >> (9 - ks));
}

enum SlotKind : int { K_NONE = 0, K_PAWN = 1, K_KNIGHT = 2, K_SLIDER = 3, K_KING = 4 };

struct Slot {
    int kind;
    uint64_t tgt;      // destination squares
    uint64_t epm;      // pawn: destinations that are e.p. captures
    uint32_t aux;      // slider: direction mask; king: castle bits
};

KV_DEV int slot_count(const Slot& s) {
    return s.kind == K_NONE ? 0 : popc64(s.tgt) + (s.kind == K_KING ? popc32(s.aux) : 0);
}

// Write the slot's moves to mv[off..] in the reference's emission order.  Returns the new offset.
KV_DEV int slot_emit(const Tables& T, const Pos& p, const Slot& sl, int s, uint16_t* mv, int off) {
    auto put = [&](int from, int to, int fl) {
        if (off < MAX_MOVES) mv[off] = (uint16_t)(from | (to << 6) | (fl << 12));
        off++;
    };
    if (sl.kind == K_PAWN) {
        // push1, push2, capture dc=-1, capture dc=+1 (core/chessEngine.py:458-472); promotion by rank (:706-710)
        const int fwd = p.wtm ? -8 : 8;
        const int last = p.wtm ? 0 : 7;
        const int cand[4] = {s + fwd, s + 2 * fwd, s + fwd - 1, s + fwd + 1};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = cand[k];
            if (t >= 0 && t < 64 && (sl.tgt & bit(t))) {
                int fl = ((sl.epm & bit(t)) ? MF_EP : 0) | (((t >> 3) == last) ? MF_PROMO : 0);
                put(s, t, fl);
            }
        }
    } else if (sl.kind == K_KNIGHT) {
        const int dl[8] = {-17, -10, -15, -6, 6, 15, 10, 17};   // :501-502 order
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int t = s + dl[k];
            if (t >= 0 && t < 64 && (sl.tgt & bit(t))) put(s, t, 0);
        }
    } else if (sl.kind == K_SLIDER) {
        for (int d = 0; d < 8; d++) {
            if (!((sl.aux >> d) & 1)) continue;
            uint64_t m = sl.tgt & T.ray[d][s];
            while (m) {
                const int t = first_on_ray(m, d);
                m ^= bit(t);
                put(s, t, 0);
            }
        }
    } else if (sl.kind == K_KING) {
        const int dl[8] = {-9, -8, -7, -1, 1, 7, 8, 9};          // :544-546 order
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int t = s + dl[k];
            if (t >= 0 && t < 64 && (sl.tgt & bit(t))) put(s, t, 0);
        }
        const int home = p.wtm ? 60 : 4;
        if (sl.aux & 1) put(home, home + 2, MF_CASTLE);
        if (sl.aux & 2) put(home, home - 2, MF_CASTLE);
    }
    return off;
}

struct GenOut {
    int n;       // number of legal moves (may exceed MAX_MOVES; only MAX_MOVES are stored)
    int flags;   // RF_*
};

// What getValidMoves decides before any piece moves are generated.
struct GenState {
    Pos p;
    Agg g;
    uint64_t valid, pinned;
    int mode;    // 0 no check, 1 one check, 2 two or more (king moves only)
    int flags;
    int n_slots; // pieces that generate (mode 2: the king square alone)
};

// getValidMoves (core/chessEngine.py:277-321), part 1: pins and checks from the king location variable.
//   L       the board line as the group's lanes hold it; updated in place when the getKingMoves restore quirk
//           (:564, stale king location) rewrites the board (flags & RF_STATE_MUTATED)
template <int W>
KV_DEV void movegen_prepare(const Tables& T, int lane, Line<W>& L, GenState& S) {
    Pos& p = S.p;
    Agg& g = S.g;
    const int q = sub_q<W>(lane);
    load_pos<W>(L, p, lane);
    make_agg(p, g);
    int flags = 0;

    // ---- checkForPinsAndChecks: lanes 0-7 one ray each -------------------------------------------------
    uint64_t pinb = 0, validm = 0;
    bool chk = false;
    if (q < 8) {
        const int d = q;
        const uint64_t r = T.ray[d][p.kloc];
        const uint64_t b = r & g.occ;
        if (b) {
            const int f = first_on_ray(b, d);
            const uint64_t fb = bit(f);
            const uint64_t sliders = d < 4 ? g.eRQ : g.eBQ;
            if (g.own & fb) {
                const uint64_t b2 = T.ray[d][f] & g.occ;
                if (b2) {
                    const int s2 = first_on_ray(b2, d);
                    if (sliders & bit(s2)) pinb = fb;
                }
            } else {
                bool m = (sliders & fb) != 0;
                const bool pawn_dir = p.wtm ? (d == D_NW || d == D_NE) : (d == D_SW || d == D_SE);
                if (!m && pawn_dir && (g.eP & fb) && (T.king[p.kloc] & fb)) m = true;
                if (m) {
                    chk = true;
                    validm = r ^ T.ray[d][f];
                }
            }
        }
    }
    const uint32_t chkb = sub_ballot<W>(chk, lane);
    const uint64_t kn = T.knight7[p.kloc] & g.eN;   // 7-entry table: (-2,+1) is missing
    const int nchecks = popc32(chkb) + popc64(kn);
    const uint64_t pinned = sub_or64<W>(pinb, lane);
    const int src = chkb ? ffs32(chkb) - 1 : 0;
    uint64_t valid = sub_shfl64<W>(validm, src, lane);
    if (kn) valid = kn;
    const int mode = nchecks == 0 ? 0 : (nchecks == 1 ? 1 : 2);
    if (nchecks) flags |= RF_E3_CHECK;

    // ---- >= 2 checks: getKingMoves(kingRow, kingCol) on whatever stands there (:315).  Its restore step
    // (:564) leaves a king on the square once any step qualified; apply that rewrite up front. -----------
    if (mode == 2) {
        const uint64_t kb = bit(p.kloc);
        if ((T.king[p.kloc] & ~g.own) && !(p.o[T_K] & kb)) {
#pragma unroll
            for (int i = 0; i < 6; i++) {
                p.o[i] &= ~kb;
                p.e[i] &= ~kb;
            }
            p.o[T_K] |= kb;
            make_agg(p, g);
            flags |= RF_STATE_MUTATED;
            // write the rewritten board back into the bitboard words (0-11) the lanes of the group hold
#pragma unroll
            for (int j = 0; j < Line<W>::NW; j++) {
                const int idx = q + j * Line<W>::WL;
                if (idx < 12) {
                    L.w[j] &= ~kb;
                    if (idx == (p.wtm ? 0 : 6)) L.w[j] |= kb;
                }
            }
        }
    }
    S.valid = valid;
    S.pinned = pinned;
    S.mode = mode;
    S.flags = flags;
    S.n_slots = mode == 2 ? 1 : popc64(g.own);
}

// Destination set of the non-king piece of the side to move on square s (:433-441, :604-630 pin filter, :296-311
// check filter).
KV_DEV void piece_slot(const Tables& T, const GenState& S, int s, Slot& o) {
    const Pos& p = S.p;
    const Agg& g = S.g;
    const uint64_t sb = bit(s);
    const bool is_pinned = (S.pinned & sb) != 0;
    int pd = 0;
    if (is_pinned)
        for (int d = 0; d < 8; d++)
            if (T.ray[d][p.kloc] & sb) pd = d;
    uint64_t tgt = 0;
    if (p.o[T_P] & sb) {
        o.kind = K_PAWN;
        const int r = s >> 3, c = s & 7;
        const int dr = p.wtm ? -1 : 1;
        const int fwd = dr * 8;
        const int r1 = r + dr;
        if (r1 >= 0 && r1 < 8) {
            if (!is_pinned || pd == (p.wtm ? D_N : D_S)) {
                if (!(g.occ & bit(s + fwd))) {
                    tgt |= bit(s + fwd);
                    if (r == (p.wtm ? 6 : 1) && !(g.occ & bit(s + 2 * fwd))) tgt |= bit(s + 2 * fwd);
                }
            }
            if (c > 0 && (!is_pinned || pd == (p.wtm ? D_NW : D_SW))) {
                const int t = s + fwd - 1;
                if (g.opp & bit(t)) tgt |= bit(t);
                else if (t == p.ep) { tgt |= bit(t); o.epm |= bit(t); }
            }
            if (c < 7 && (!is_pinned || pd == (p.wtm ? D_NE : D_SE))) {
                const int t = s + fwd + 1;
                if (g.opp & bit(t)) tgt |= bit(t);
                else if (t == p.ep) { tgt |= bit(t); o.epm |= bit(t); }
            }
        }
    } else if (p.o[T_N] & sb) {
        o.kind = K_KNIGHT;
        if (!is_pinned) tgt = T.knight[s] & ~g.own;
    } else {
        o.kind = K_SLIDER;
        o.aux = (p.o[T_R] & sb) ? 0x0Fu : ((p.o[T_B] & sb) ? 0xF0u : 0xFFu);
        if (o.aux & 0x0F) tgt |= rook_att(T, s, g.occ);
        if (o.aux & 0xF0) tgt |= bishop_att(T, s, g.occ);
        tgt &= ~g.own;
        if (is_pinned) tgt &= T.ray[pd][s] | T.ray[dir_opp(pd)][s];
    }
    if (S.mode == 1) tgt &= S.valid;
    o.tgt = tgt;
    o.epm &= tgt;
}

// getValidMoves, part 2: the pieces in square order, W per round; per_round(slot, square) is called by every lane of
// the warp once per round (lanes without a piece pass a K_NONE slot), so it may use collectives.
template <int W, class F>
KV_DEV void movegen_rounds(const Tables& T, int lane, GenState& S, F&& per_round) {
    const Pos& p = S.p;
    const int q = sub_q<W>(lane);
    for (int r = 0;; r++) {
        const int i = r * W + q;
        const bool act = i < S.n_slots;
        // every board of the warp takes the same number of rounds (the collectives below need all 32 lanes)
        if (W == 32) {
            if (r * W >= S.n_slots) break;
        } else if (ballot(r * W < S.n_slots) == 0) {
            break;
        }
        Slot sl;
        sl.kind = K_NONE;
        sl.tgt = sl.epm = 0;
        sl.aux = 0;
        const int s = act ? (S.mode == 2 ? p.kloc : nth_set64(S.g.own, i)) : 0;
        // king(s) of this round: one collective king_phase each (one on any reachable position)
        const bool king_slot = act && (S.mode == 2 || (p.o[T_K] & bit(s)));
        uint32_t kb = sub_ballot<W>(king_slot, lane);
        while (W == 32 ? kb != 0 : ballot(kb != 0) != 0) {
            const bool has = kb != 0;
            const int src = has ? ffs32(kb) - 1 : 0;
            const int ks = sub_shfl32<W>(s, src, lane);
            kb &= kb - 1;
            const KingOut ko = king_phase<W>(T, lane, p, S.g, ks, S.mode, S.valid, S.pinned);
            if (has && q == src) {
                sl.kind = K_KING;
                sl.tgt = king_steps_to_mask(ks, ko.steps);
                sl.aux = ko.castle;
            }
        }
        if (act && !king_slot) piece_slot(T, S, s, sl);
        per_round(sl, s);
    }
}

// checkForEndConditions (:632-651, repetition excluded) + isDraw's only-kings clause, given the move count n
KV_DEV int movegen_end_flags(const Tables& T, const GenState& S, int n) {
    const Pos& p = S.p;
    int flags = S.flags;
    if (n == 0) {
        const bool ic = attacked(T, S.g, p.kloc, p.wtm, p.ep, p.moved, p.akloc);   // inCheck(), :388-394
        flags |= ic ? RF_CHECKMATE : RF_STALEMATE;
    } else if (p.clock >= 100) {
        flags |= RF_DRAW50;
    }
    if (S.g.occ == (p.o[T_K] | p.e[T_K])) flags |= RF_ONLY_KINGS;   // isDraw(), :21-33 (the only satisfiable clause)
    if (n > MAX_MOVES) flags |= RF_OVERFLOW;
    return flags;
}

// getValidMoves with the ordered move list written to mv (the board's MAX_MOVES u16 of shared memory:
// from | to<<6 | ep<<12 | castle<<13 | promo<<14).
template <int W>
KV_DEV GenOut movegen_sub(const Tables& T, int lane, Line<W>& L, uint16_t* mv) {
    GenState S;
    movegen_prepare<W>(T, lane, L, S);
    int n = 0;
    movegen_rounds<W>(T, lane, S, [&](const Slot& sl, int s) {
        const int c = slot_count(sl);
        const int incl = sub_incl_scan<W>(c, lane);
        slot_emit(T, S.p, sl, s, mv, n + incl - c);
        n += sub_shfl32<W>(incl, W - 1, lane);
    });
    syncwarp();
    GenOut out;
    out.n = n;
    out.flags = movegen_end_flags(T, S, n);
    return out;
}

// Bulk count without laying the list out (perft leaves): n, and in `cats` five 12-bit fields
// n | captures<<12 | ep<<24 | castles<<36 | promos<<48
// (a capture = destination occupied or e.p., Move.pieceCaptured != "--", core/chessEngine.py:699-703).
template <int W>
KV_DEV GenOut movegen_count_sub(const Tables& T, int lane, Line<W>& L, uint64_t& cats) {
    GenState S;
    movegen_prepare<W>(T, lane, L, S);
    const uint64_t last = S.p.wtm ? 0xFFull : 0xFFull << 56;
    // five 12-bit fields in one word (each <= 256 per lane and <= MAX_MOVES-ish per board): n, captures, ep, castles, promos
    uint64_t acc = 0;
    movegen_rounds<W>(T, lane, S, [&](const Slot& sl, int) {
        if (sl.kind == K_NONE) return;
        const uint64_t ncast = sl.kind == K_KING ? (uint64_t)popc32(sl.aux) : 0;
        const uint64_t nep = (uint64_t)popc64(sl.epm);
        const uint64_t cap = (uint64_t)popc64(sl.tgt & S.g.opp) + nep;
        const uint64_t pro = sl.kind == K_PAWN ? (uint64_t)popc64(sl.tgt & last) : 0;
        acc += ((uint64_t)popc64(sl.tgt) + ncast) | (cap << 12) | (nep << 24) | (ncast << 36) | (pro << 48);
    });
    acc = sub_sum64<W>(acc, lane);
    const int n = (int)(acc & 0xFFF);
    cats = acc;
    GenOut out;
    out.n = n;
    out.flags = movegen_end_flags(T, S, n);
    return out;
}

// one board per warp (the tree-search kernels)
KV_DEV GenOut movegen_warp(const Tables& T, int lane, uint64_t& w, uint16_t* mv) {
    Line<32> L;
    L.w[0] = w;
    const GenOut g = movegen_sub<32>(T, lane, L, mv);
    w = L.w[0];
    return g;
}

// makeMove (core/chessEngine.py:127-197), no legality check, mailbox write order preserved.
// The lanes of the board's group hold the line (Line<W>): bitboards are words 0-11, the meta word is 12.
// Convergent for any mix of moves in a warp (every ballot is reached by all lanes).
template <int W>
KV_DEV void make_move_sub(int lane, Line<W>& L, int mvw, int promo_type) {
    constexpr int WL = Line<W>::WL, NW = Line<W>::NW;
    const int q = sub_q<W>(lane);
    const int from = mvw & 63, to = (mvw >> 6) & 63, fl = (mvw >> 12) & 7;
    const uint64_t fb = bit(from), tb = bit(to);
    // which bitboard holds a square: one ballot per word the lanes hold (lowest bitboard index wins, as a mailbox would)
    auto holder = [&](uint64_t sq_bit, bool enable) {
        int idx = -1;
#pragma unroll
        for (int j = NW - 1; j >= 0; j--) {
            const uint32_t b = sub_ballot<W>(enable && (q + j * WL) < 12 && (L.w[j] & sq_bit), lane);
            if (b) idx = j * WL + ffs32(b) - 1;
        }
        return idx;
    };
    const int pm = holder(fb, true);
    const bool has_t = holder(tb, true) >= 0;
    const bool captured = (fl & MF_EP) || has_t;
    const int sr = from >> 3, sc = from & 7, er = to >> 3, ec = to & 7;
    int rs = -1, rd = -1;               // rook hop, :156-164
    if (fl & MF_CASTLE) {
        if (ec - sc == 2) {
            if (ec + 1 < 8) { rs = er * 8 + ec + 1; rd = er * 8 + ec - 1; }
        } else if (ec - 2 >= 0 && ec + 1 < 8) {
            rs = er * 8 + ec - 2; rd = er * 8 + ec + 1;
        }
    }
#pragma unroll
    for (int j = 0; j < NW; j++) {
        const int idx = q + j * WL;
        if (idx < 12) {
            uint64_t w = L.w[j];
            w &= ~fb;                       // board[start] = "--"
            w &= ~tb;                       // board[end] = pieceMoved
            if (idx == pm) w |= tb;
            if (fl & MF_EP) w &= ~bit(sr * 8 + ec);   // :152-153
            L.w[j] = w;
        }
    }
    const int rp = holder(bit(rs & 63), rs >= 0);   // the rook hop reads the board after the king has moved
    if (rs >= 0) {
#pragma unroll
        for (int j = 0; j < NW; j++) {
            const int idx = q + j * WL;
            if (idx < 12) {
                uint64_t w = L.w[j];
                w &= ~bit(rd);
                if (idx == rp) w |= bit(rd);
                w &= ~bit(rs);
                L.w[j] = w;
            }
        }
    }
    if (fl & MF_PROMO) {                // :190-191, promotionChoice defaults to 'Q'
        const int pp = ((pm >= 0 && pm < 6) ? 0 : 6) + promo_type;
#pragma unroll
        for (int j = 0; j < NW; j++) {
            const int idx = q + j * WL;
            if (idx < 12) {
                L.w[j] &= ~tb;
                if (idx == pp) L.w[j] |= tb;
            }
        }
    }
    constexpr int MJ = 12 / WL, MQ = 12 % WL;   // where the meta word lives
    if (q == MQ) {
        uint64_t m = L.w[MJ];
        int moved = (int)((m >> 1) & 63);
        int wk = (int)((m >> 16) & 63), bk = (int)((m >> 24) & 63);
        int clock = (int)((m >> 32) & 0xFFFF);
        const bool wtm = m & 1;
        if (pm == 0) { moved |= F_WK; wk = to; }
        else if (pm == 6) { moved |= F_BK; bk = to; }
        else if (pm == 2) { if (from == 56) moved |= F_WRQ; else if (from == 63) moved |= F_WRK; }
        else if (pm == 8) { if (from == 0) moved |= F_BRQ; else if (from == 7) moved |= F_BRK; }
        int ep = EP_NONE;
        if ((pm == 5 || pm == 11) && (sr - er == 2 || er - sr == 2)) ep = ((sr + er) >> 1) * 8 + sc;   // :169-173
        clock = captured ? 0 : (clock + 1 > 0xFFFF ? 0xFFFF : clock + 1);   // :178 resets on captures only
        L.w[MJ] = (m & 0xFFFF000000000000ull) | (uint64_t)(wtm ? 0 : 1) | ((uint64_t)moved << 1) | ((uint64_t)ep << 8) |
                  ((uint64_t)wk << 16) | ((uint64_t)bk << 24) | ((uint64_t)clock << 32);
    }
}
KV_DEV uint64_t make_move_warp(int lane, uint64_t w, int mvw, int promo_type) {
    Line<32> L;
    L.w[0] = w;
    make_move_sub<32>(lane, L, mvw, promo_type);
    return L.w[0];
}

// squareUnderAttack(r, c) for all 64 squares of one board (core/chessEngine.py:400-415): bit r*8+c of the result.
KV_DEV uint64_t attacked_mask_warp(const Tables& T, int lane, uint64_t w) {
    Pos p;
    Line<32> L;
    L.w[0] = w;
    load_pos<32>(L, p, lane);
    Agg g;
    make_agg(p, g);
    const bool a0 = attacked(T, g, 2 * lane, p.wtm, p.ep, p.moved, p.akloc);
    const bool a1 = attacked(T, g, 2 * lane + 1, p.wtm, p.ep, p.moved, p.akloc);
    const uint32_t b0 = ballot(a0), b1 = ballot(a1);
    uint64_t m = 0;
#pragma unroll
    for (int l = 0; l < 32; l++) m |= ((uint64_t)((b0 >> l) & 1) << (2 * l)) | ((uint64_t)((b1 >> l) & 1) << (2 * l + 1));
    return m;
}

// ---- perft (kv_perft): one frontier board per group of W lanes ---------------------------------------------
KV_DEV uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
// order digest of one move list under path hash `path`: sum_k mix64(path + (k+1)*G + (mv_k << 32))
template <int W>
KV_DEV uint64_t list_digest(const uint16_t* mv, int n, uint64_t path, int lane) {
    uint64_t h = 0;
    for (int k = sub_q<W>(lane); k < n; k += W)
        h += mix64(path + (uint64_t)(k + 1) * 0x9E3779B97F4A7C15ull + ((uint64_t)mv[k] << 32));
    return sub_sum64<W>(h, lane);
}
KV_DEV uint64_t child_path(uint64_t path, int k) { return mix64(path ^ ((uint64_t)(k + 1) * 0xD6E8FEB86659FD93ull)); }

// lane q of the group owns accumulator q: nodes, captures, ep, castles, promos, digest, movegen calls
KV_DEV void perft_acc_flush(uint64_t& accv, int root, uint64_t* out, int q) {
    if (root >= 0 && q < 7 && accv) atomic_add_u64(out + (size_t)root * 8 + q, accv);
    accv = 0;
}

// Visit one frontier board per group (valid = the group has one; a group without still runs the collectives, on an
// empty line).  LEAF: bulk-count its move list.  Otherwise: write every child board to `next` (slots claimed with one
// atomic add per parent; the frontier is an unordered multiset and every output is an order-independent sum).
// w[13] = root id, w[14] = path hash.
template <int W, bool LEAF, bool DIGEST = true>
KV_DEV void perft_visit_sub(const Tables& T, int lane, Line<W> L, bool valid, uint16_t* mv, uint64_t& accv, int& acc_root,
                            uint64_t* next, uint32_t* next_count, uint64_t* out) {
    constexpr int WL = Line<W>::WL, NW = Line<W>::NW;
    const int q = sub_q<W>(lane);
    const int root_w = (int)(uint32_t)line_word<W>(L, 13, lane);   // (a collective: outside the conditional)
    const int root = valid ? root_w : acc_root;
    const uint64_t path = line_word<W>(L, 14, lane);
    if (root != acc_root) {
        perft_acc_flush(accv, acc_root, out, q);
        acc_root = root;
    }
    if (LEAF && !DIGEST) {   // count-only leaves: no ordered list, no digest
        uint64_t cats = 0;
        const GenOut g = movegen_count_sub<W>(T, lane, L, cats);
        if (valid) {
            if (q == 0) accv += (unsigned)(g.n < MAX_MOVES ? g.n : MAX_MOVES);
            if (q >= 1 && q <= 4) accv += (cats >> (12 * q)) & 0xFFF;
            if (q == 6) accv += 1;
        }
        return;
    }
    // w reflects the :564 rewrite afterwards, as the reference's makeMove would see it
    const GenOut g = movegen_sub<W>(T, lane, L, mv);
    const int n = valid ? (g.n < MAX_MOVES ? g.n : MAX_MOVES) : 0;
    const uint64_t dig = list_digest<W>(mv, n, path, lane);
    if (LEAF) {
        uint64_t mine = 0;
#pragma unroll
        for (int j = 0; j < NW; j++)
            if (q + j * WL < 12) mine |= L.w[j];
        const uint64_t occ = sub_or64<W>(mine, lane);
        uint64_t cats = 0;   // captures | ep<<16 | castles<<32 | promos<<48 (each <= 256)
        for (int k = q; k < n; k += W) {
            const int x = mv[k];
            const int fl = x >> 12;
            const uint64_t cap = ((fl & 1) || ((occ >> ((x >> 6) & 63)) & 1)) ? 1 : 0;
            cats += cap | ((uint64_t)(fl & 1) << 16) | ((uint64_t)((fl >> 1) & 1) << 32) |
                    ((uint64_t)((fl >> 2) & 1) << 48);
        }
        cats = sub_sum64<W>(cats, lane);
        if (q == 0) accv += (unsigned)n;
        if (q >= 1 && q <= 4) accv += (cats >> (16 * (q - 1))) & 0xFFFF;
    } else {
        uint32_t base = 0;
        if (q == 0 && n) base = atomic_add_u32(next_count, (uint32_t)n);
        base = (uint32_t)sub_shfl32<W>((int)base, 0, lane);
        int nmax = n;                       // every group makes as many moves as the busiest one of the warp
        if (W < 32) {
#pragma unroll
            for (int m = 16; m >= W; m >>= 1) {
                const int o = shfl_xor32(nmax, m, lane);
                nmax = o > nmax ? o : nmax;
            }
        }
        for (int k = 0; k < nmax; k++) {
            const bool live = k < n;
            Line<W> C = L;
            make_move_sub<W>(lane, C, live ? mv[k] : 0, T_Q);
#pragma unroll
            for (int j = 0; j < NW; j++) {
                const int idx = q + j * WL;
                if (idx == 14) C.w[j] = child_path(path, k);
                if (live && idx < LINE_WORDS) next[(size_t)(base + k) * LINE_WORDS + idx] = C.w[j];
            }
        }
    }
    if (valid) {
        if (q == 5 && DIGEST) accv += dig;
        if (q == 6) accv += 1;
    }
    syncwarp();
}

}  // namespace kv
