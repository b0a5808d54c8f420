// kv_tower_order.h — task order of the whole-tower launch (tower_umma2_kernel, kv_net.cu).  Plain C++ so that the CPU
// tests can check it (tests/test_host_logic.py through the emulator library): every (layer, board tile, channel tile)
// exactly once, and a tile always after the tiles it reads.
#pragma once
#ifdef __CUDACC__
#define KV_TO_HD __host__ __device__
#else
#define KV_TO_HD
#endif

namespace kvn {

// Task order.  Layer-major over ALL boards streams every layer's activations through HBM (a layer of 4 096 boards is
// 268 MB, twice the L2).  Depth-first in chunks of Mc board tiles — all layers of chunk 0, then chunk 1, ... — keeps the
// three ping-pong buffers of a chunk (Mc x 768 KB) in the 126 MB L2: a layer's output is read back from L2 and is
// overwritten there before it is ever evicted.  Inside a chunk the order is layer-major, so a task's inputs were
// finished Mc * NT tasks (>= 2 rounds of the 74 CTA pairs) ago.  The chunks are equal up to one tile (no small last
// chunk whose layers would wait for each other).
struct TowerOrder {
    int M, NT, n_layers, n_chunks, Mc, big;   // the first `big` chunks have Mc tiles, the others Mc - 1
    int total;
    KV_TO_HD void init(int M_, int NT_, int n_layers_, int chunk_tiles) {
        M = M_; NT = NT_; n_layers = n_layers_;
        n_chunks = (chunk_tiles > 0 && M > 0) ? (M / chunk_tiles > 1 ? M / chunk_tiles : 1) : 1;
        Mc = (M + n_chunks - 1) / n_chunks;
        big = M - (Mc - 1) * n_chunks;
        total = n_layers * M * NT;
    }
    KV_TO_HD void decode(int t, int& l, int& m_tile, int& n_tile) const {
        const int full = n_layers * Mc * NT, tb = big * full;
        int c, mc, m0;
        if (t < tb) {
            c = t / full; t -= c * full; mc = Mc; m0 = c * Mc;
        } else {
            const int small = n_layers * (Mc - 1) * NT;
            t -= tb; c = t / small; t -= c * small; mc = Mc - 1; m0 = big * Mc + c * (Mc - 1);
        }
        l = t / (mc * NT);
        const int r = t - l * mc * NT;
        m_tile = m0 + r / NT;
        n_tile = r - (r / NT) * NT;
    }
};

// Hybrid launch (4-CTA clusters + CTA pairs, kv_net.cu): the slice of the M board tiles a kernel computes.  part 0: all;
// part 1: [0, Ma); part 2: [Ma, M); Ma = M * num / den rounded down to even (the clusters take board tiles in pairs).
KV_TO_HD inline void tower_slice_tiles(int part, int num, int den, int M, int& m_off, int& m_cnt) {
    m_off = 0;
    m_cnt = M;
    if (part) {
        const int Ma = (int)(((long long)M * num / den) & ~1ll);
        if (part == 1) m_cnt = Ma;
        else {
            m_off = Ma;
            m_cnt = M - Ma;
        }
    }
}

}  // namespace kvn
