// kv_tables.cuh — attack tables for the bitboard rules kernels (5.5 KB, staged in shared memory).
//
// Square index = row*8 + col with row 0 = rank 8, exactly the reference's board[row][col]
// (core/chessEngine.py:39-47).  Direction order = the reference's piece generators: rook
// (-1,0),(1,0),(0,-1),(0,1) (core/chessEngine.py:478), then bishop (-1,-1),(-1,1),(1,-1),(1,1) (:517).
#pragma once
#include <cstdint>

namespace kv {

enum Dir : int { D_N = 0, D_S = 1, D_W = 2, D_E = 3, D_NW = 4, D_NE = 5, D_SW = 6, D_SE = 7 };

struct Tables {
    uint64_t ray[8][64];    // squares strictly beyond s in direction d, to the board edge
    uint64_t knight[64];    // getKnightMoves, core/chessEngine.py:500-512 (all 8 offsets)
    uint64_t knight7[64];   // checkForPinsAndChecks' 7-entry table (no (-2,+1)), :373-374
    uint64_t king[64];      // getKingMoves' 8 steps, :544-546
};

constexpr int kDirR[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
constexpr int kDirC[8] = {0, 0, -1, 1, -1, 1, -1, 1};

constexpr Tables make_tables() {
    Tables t{};
    constexpr int kn[8][2] = {{-2, -1}, {-1, -2}, {-2, 1}, {-1, 2}, {1, -2}, {2, -1}, {1, 2}, {2, 1}};
    for (int s = 0; s < 64; s++) {
        int r = s >> 3, c = s & 7;
        for (int d = 0; d < 8; d++) {
            uint64_t m = 0;
            for (int i = 1; i < 8; i++) {
                int er = r + kDirR[d] * i, ec = c + kDirC[d] * i;
                if (er < 0 || er >= 8 || ec < 0 || ec >= 8) break;
                m |= 1ull << (er * 8 + ec);
            }
            t.ray[d][s] = m;
        }
        uint64_t n8 = 0, n7 = 0, k8 = 0;
        for (int k = 0; k < 8; k++) {
            int er = r + kn[k][0], ec = c + kn[k][1];
            if (er >= 0 && er < 8 && ec >= 0 && ec < 8) {
                n8 |= 1ull << (er * 8 + ec);
                if (!(kn[k][0] == -2 && kn[k][1] == 1)) n7 |= 1ull << (er * 8 + ec);
            }
        }
        for (int dr = -1; dr <= 1; dr++)
            for (int dc = -1; dc <= 1; dc++) {
                if (!dr && !dc) continue;
                int er = r + dr, ec = c + dc;
                if (er >= 0 && er < 8 && ec >= 0 && ec < 8) k8 |= 1ull << (er * 8 + ec);
            }
        t.knight[s] = n8;
        t.knight7[s] = n7;
        t.king[s] = k8;
    }
    return t;
}

constexpr int kTableWords = sizeof(Tables) / 8;

}  // namespace kv
