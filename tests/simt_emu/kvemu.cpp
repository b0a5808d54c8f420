// kvemu.cpp — CPU lock-step emulator for the warp-per-board integer kernels (TEST INFRASTRUCTURE ONLY).
//
// Compiles knightvision_b200/csrc/kv_rules.cuh (and kv_mcts.cuh) with KV_HOST_EMU: the 32 lanes of a
// warp run as 32 ucontext fibers; every warp collective (shfl / ballot / syncwarp) is a rendezvous —
// a lane posts its value, yields to the scheduler, and reads its peers' values once all 32 have posted.
// This lets `pytest -m "not gpu"` check the *kernel source* against the oracle on a box without a GPU.
// It is never part of the product library, and GPU tests never use it.
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace kvemu {
static ucontext_t g_sched;
static ucontext_t g_fib[32];
static char* g_stack[32];
static int g_cur = -1;
static bool g_done[32];
static uint64_t g_slot[2][32];
static int g_parity[32];
static long g_ncoll[32];
static std::function<void(int)>* g_body;
constexpr size_t kStack = 256 * 1024;

static void yield_to_sched() { swapcontext(&g_fib[g_cur], &g_sched); }

uint64_t collective_shfl(uint64_t v, int src) {
    const int me = g_cur, ph = g_parity[me];
    g_slot[ph][me] = v;
    g_ncoll[me]++;
    yield_to_sched();
    g_parity[me] ^= 1;
    return g_slot[ph][src & 31];
}
uint32_t collective_ballot(bool p) {
    const int me = g_cur, ph = g_parity[me];
    g_slot[ph][me] = p ? 1 : 0;
    g_ncoll[me]++;
    yield_to_sched();
    g_parity[me] ^= 1;
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= (uint32_t)(g_slot[ph][i] & 1) << i;
    return m;
}
void collective_sync() { (void)collective_shfl(0, 0); }
int lane_id() { return g_cur; }

static void trampoline() {
    (*g_body)(g_cur);
    g_done[g_cur] = true;
    yield_to_sched();
}

// Run body(lane) for lanes 0..31 in lock step.
void run_warp(std::function<void(int)> body) {
    g_body = &body;
    for (int i = 0; i < 32; i++) {
        if (!g_stack[i]) g_stack[i] = (char*)malloc(kStack);
        getcontext(&g_fib[i]);
        g_fib[i].uc_stack.ss_sp = g_stack[i];
        g_fib[i].uc_stack.ss_size = kStack;
        g_fib[i].uc_link = &g_sched;
        makecontext(&g_fib[i], trampoline, 0);
        g_done[i] = false;
        g_parity[i] = 0;
        g_ncoll[i] = 0;
    }
    for (;;) {
        int alive = 0;
        for (int i = 0; i < 32; i++) {
            if (g_done[i]) continue;
            g_cur = i;
            swapcontext(&g_sched, &g_fib[i]);
            if (!g_done[i]) alive++;
        }
        if (alive == 0) break;
        if (alive != 32) {
            // a lane left while others still wait at a collective: the kernel is not convergent
            bool any_done = false;
            for (int i = 0; i < 32; i++) any_done |= g_done[i];
            if (any_done) {
                fprintf(stderr, "kvemu: divergent collective (lanes exited early)\n");
                abort();
            }
        }
    }
    for (int i = 1; i < 32; i++)
        if (g_ncoll[i] != g_ncoll[0]) {
            fprintf(stderr, "kvemu: lanes executed different collective counts\n");
            abort();
        }
    g_cur = -1;
}
}  // namespace kvemu

#define KV_HOST_EMU 1
#include "../../knightvision_b200/csrc/kv_rules.cuh"
#include "../../knightvision_b200/csrc/kv_mcts.cuh"
#include "../../knightvision_b200/csrc/kv_tower_order.h"

static const kv::Tables g_tables = kv::make_tables();

// Lanes per board in the rules entry points below: 32 (one board per warp, what the tree-search kernels use), 16 or 8
// (two / four boards per warp, what kv_rules.cu's kernels can be built for).  Trailing boards that do not fill a warp
// leave groups without a board.
static int g_emu_w = 32;

template <int W>
static kv::Line<W> emu_load_line(const uint64_t* line, int q, bool valid) {
    kv::Line<W> L;
    for (int j = 0; j < kv::Line<W>::NW; j++) {
        const int idx = q + j * kv::Line<W>::WL;
        L.w[j] = (valid && idx < 16) ? line[idx] : 0;
    }
    return L;
}
template <int W>
static void emu_store_line(uint64_t* line, int q, const kv::Line<W>& L) {
    for (int j = 0; j < kv::Line<W>::NW; j++) {
        const int idx = q + j * kv::Line<W>::WL;
        if (idx < 16) line[idx] = L.w[j];
    }
}

template <int W>
static void emu_movegen_t(uint64_t* lines, int n, uint16_t* moves, int stride, int32_t* counts, int32_t* flags) {
    constexpr int NB = 32 / W;
    for (int i0 = 0; i0 < n; i0 += NB) {
        uint16_t mv[NB][kv::MAX_MOVES];
        memset(mv, 0, sizeof(mv));
        kv::GenOut out[32];
        uint64_t neww[NB][16];
        kvemu::run_warp([&](int lane) {
            const int grp = lane / W, q = lane % W, i = i0 + grp;
            kv::Line<W> L = emu_load_line<W>(lines + 16 * (size_t)(i < n ? i : 0), q, i < n);
            out[lane] = kv::movegen_sub<W>(g_tables, lane, L, mv[grp]);
            emu_store_line<W>(neww[grp], q, L);
        });
        for (int grp = 0; grp < NB && i0 + grp < n; grp++) {
            const int i = i0 + grp;
            uint64_t* line = lines + 16 * (size_t)i;
            for (int l = 1; l < W; l++)
                if (out[grp * W + l].n != out[grp * W].n || out[grp * W + l].flags != out[grp * W].flags) {
                    fprintf(stderr, "kvemu: non-uniform movegen result\n");
                    abort();
                }
            const kv::GenOut o = out[grp * W];
            if (o.flags & kv::RF_STATE_MUTATED)
                for (int l = 0; l < 16; l++) line[l] = neww[grp][l];
            counts[i] = o.n;
            flags[i] = o.flags;
            for (int k = 0; k < o.n && k < stride && k < kv::MAX_MOVES; k++) moves[(size_t)i * stride + k] = mv[grp][k];
        }
    }
}

template <int W>
static void emu_make_moves_t(uint64_t* lines, int n, const uint16_t* mv) {
    constexpr int NB = 32 / W;
    for (int i0 = 0; i0 < n; i0 += NB) {
        uint64_t neww[NB][16];
        kvemu::run_warp([&](int lane) {
            const int grp = lane / W, q = lane % W, i = i0 + grp;
            const bool valid = i < n && mv[i] != 0xFFFF;
            kv::Line<W> L = emu_load_line<W>(lines + 16 * (size_t)(i < n ? i : 0), q, valid);
            kv::make_move_sub<W>(lane, L, valid ? mv[i] : 0, kv::T_Q);
            emu_store_line<W>(neww[grp], q, L);
        });
        for (int grp = 0; grp < NB && i0 + grp < n; grp++)
            if (mv[i0 + grp] != 0xFFFF)
                for (int l = 0; l < 16; l++) lines[16 * (size_t)(i0 + grp) + l] = neww[grp][l];
    }
}

// kv_perft's level loop over host memory (same chunked depth-first / breadth-first-in-chunk order)
static bool g_emu_digest = true;
template <int W>
static void emu_perft_rec(std::vector<uint64_t>& cur, int remaining, uint64_t* out) {
    constexpr int NB = 32 / W;
    const int m = (int)(cur.size() / 16);
    uint16_t mv[NB][kv::MAX_MOVES];
    const bool leaf = remaining == 1;
    std::vector<uint64_t> next(leaf ? 0 : (size_t)m * kv::MAX_MOVES * 16);
    uint32_t cnt = 0;
    for (int i0 = 0; i0 < m; i0 += NB) {
        uint64_t acc[32] = {0};
        kvemu::run_warp([&](int lane) {
            const int grp = lane / W, q = lane % W, i = i0 + grp;
            const bool valid = i < m;
            int acc_root = -1;
            kv::Line<W> L = emu_load_line<W>(cur.data() + 16 * (size_t)(valid ? i : 0), q, valid);
            if (leaf) {
                if (g_emu_digest) kv::perft_visit_sub<W, true, true>(g_tables, lane, L, valid, mv[grp], acc[lane], acc_root, nullptr, nullptr, out);
                else kv::perft_visit_sub<W, true, false>(g_tables, lane, L, valid, mv[grp], acc[lane], acc_root, nullptr, nullptr, out);
            } else {
                if (g_emu_digest) kv::perft_visit_sub<W, false, true>(g_tables, lane, L, valid, mv[grp], acc[lane], acc_root, next.data(), &cnt, out);
                else kv::perft_visit_sub<W, false, false>(g_tables, lane, L, valid, mv[grp], acc[lane], acc_root, next.data(), &cnt, out);
            }
            kv::perft_acc_flush(acc[lane], acc_root, out, q);
        });
    }
    if (leaf) return;
    next.resize((size_t)cnt * 16);
    if (cnt) emu_perft_rec<W>(next, remaining - 1, out);
}

extern "C" {

__attribute__((visibility("default"))) void kvemu_set_width(int w) { g_emu_w = (w == 16 || w == 8) ? w : 32; }

__attribute__((visibility("default"))) void kvemu_movegen(uint64_t* lines, int n, uint16_t* moves, int stride,
                                                           int32_t* counts, int32_t* flags) {
    if (g_emu_w == 8) emu_movegen_t<8>(lines, n, moves, stride, counts, flags);
    else if (g_emu_w == 16) emu_movegen_t<16>(lines, n, moves, stride, counts, flags);
    else emu_movegen_t<32>(lines, n, moves, stride, counts, flags);
}

__attribute__((visibility("default"))) void kvemu_attacked(const uint64_t* lines, int n, uint64_t* masks) {
    for (int i = 0; i < n; i++) {
        uint64_t m[32];
        kvemu::run_warp([&](int lane) {
            uint64_t w = lane < 16 ? lines[16 * (size_t)i + lane] : 0;
            m[lane] = kv::attacked_mask_warp(g_tables, lane, w);
        });
        masks[i] = m[0];
    }
}

__attribute__((visibility("default"))) void kvemu_make_moves(uint64_t* lines, int n, const uint16_t* mv) {
    if (g_emu_w == 8) emu_make_moves_t<8>(lines, n, mv);
    else if (g_emu_w == 16) emu_make_moves_t<16>(lines, n, mv);
    else emu_make_moves_t<32>(lines, n, mv);
}

__attribute__((visibility("default"))) void kvemu_perft(const uint64_t* roots, int n, int depth, uint64_t* out) {
    g_emu_digest = depth > 0;   // negative depth = counts only
    if (depth < 0) depth = -depth;
    memset(out, 0, (size_t)n * 8 * sizeof(uint64_t));
    std::vector<uint64_t> cur(roots, roots + (size_t)n * 16);
    for (int i = 0; i < n; i++) {
        cur[16 * (size_t)i + 13] = (uint32_t)i;
        cur[16 * (size_t)i + 14] = 0;
        cur[16 * (size_t)i + 15] = 0;
    }
    if (g_emu_w == 8) emu_perft_rec<8>(cur, depth, out);
    else if (g_emu_w == 16) emu_perft_rec<16>(cur, depth, out);
    else emu_perft_rec<32>(cur, depth, out);
}

}  // extern "C"

// ---- MCTS: the device driver loop of kv_mcts.cu over host memory, hash evaluator ---------------------------------
struct EmuMcts {
    kv::MctsCfg cfg;
    kv::MctsArrays A;
    int G;
    std::vector<std::vector<char>> store;
    template <class T>
    T* alloc(size_t n) {
        store.emplace_back(n * sizeof(T), 0);
        return reinterpret_cast<T*>(store.back().data());
    }
};

// game-loop rules / script applied to the next contexts (kvemu_set_rules, kvemu_set_script); defaults = kv_mcts_create_k's
static float g_emu_resign_thr = -0.7f;
static int g_emu_resign_min = 15, g_emu_root_mix = -1;
static const uint16_t* g_emu_script_moves = nullptr;
static const float* g_emu_script_vals = nullptr;
static int g_emu_script_stride = 0;

static EmuMcts* emu_mcts_new(int G, int sims, int edges_per_node, int max_plies, int temp_plies, float c_puct,
                             float dir_alpha, float dir_eps, uint64_t seed, int cache_log2 = 0, int inflight = 1) {
    EmuMcts* m = new EmuMcts();
    m->G = G;
    if (inflight < 1) inflight = 1;
    m->cfg.inflight = inflight;
    kv::MctsCfg& c = m->cfg;
    c.sims = sims;
    c.node_cap = sims;
    c.edge_cap = sims * (edges_per_node > 0 ? edges_per_node : 64);
    if (c.edge_cap < kv::MAX_MOVES) c.edge_cap = kv::MAX_MOVES;
    c.temp_plies = temp_plies;
    c.max_plies = max_plies;
    c.rec_cap = max_plies;
    c.eval_mode = 0;
    c.c_puct = c_puct;
    c.dir_alpha = dir_alpha;
    c.dir_eps = dir_eps;
    c.seed = seed;
    c.resign_thr = g_emu_resign_thr;
    c.resign_min_plies = g_emu_resign_min;
    c.root_mix = g_emu_root_mix < 0 ? (sims == 1) : g_emu_root_mix;
    c.script_stride = (g_emu_script_moves || g_emu_script_vals) ? g_emu_script_stride : 0;
    kv::MctsArrays& A = m->A;
    A.script_move = c.script_stride ? g_emu_script_moves : nullptr;
    A.script_val = c.script_stride ? g_emu_script_vals : nullptr;
    const size_t g = (size_t)G, gs = g * (size_t)inflight;
    A.hdr = m->alloc<kv::GameHdr>(g);
    A.root_line = m->alloc<uint64_t>(g * 16);
    A.node_line = m->alloc<uint64_t>(g * c.node_cap * 16);
    A.node_meta = m->alloc<kv::NodeMeta>(g * c.node_cap);
    A.eP = m->alloc<float>(g * c.edge_cap);
    A.eN = m->alloc<uint32_t>(g * c.edge_cap);
    A.eW = m->alloc<float>(g * c.edge_cap);
    A.eChild = m->alloc<int>(g * c.edge_cap);
    A.eMv = m->alloc<uint16_t>(g * c.edge_cap);
    A.path_edge = m->alloc<int>(gs * (c.node_cap + 1));
    A.path_node = m->alloc<int>(gs * (c.node_cap + 1));
    A.pend_node = m->alloc<int>(gs);
    A.pend_depth = m->alloc<int>(gs);
    A.pend_kind = m->alloc<int>(gs);
    A.n_eval = m->alloc<uint32_t>(4);
    A.eval_game = m->alloc<int>(gs);
    A.eval_lines = m->alloc<uint64_t>((gs + 1) * 16);
    A.rec_line = m->alloc<uint64_t>(g * c.rec_cap * 12);
    A.rec_move = m->alloc<uint16_t>(g * c.rec_cap);
    A.eval_centry = m->alloc<int>(gs);
    A.eval_hash = m->alloc<uint64_t>(gs);
    A.n_late = m->alloc<uint32_t>(4);
    A.late_game = m->alloc<int>(gs);
    A.late_src = m->alloc<int>(gs);
    A.feat_game = m->alloc<float>(gs * kv::FEAT);
    A.feat_slot = m->alloc<float>(gs * kv::FEAT);
    A.cache = nullptr;
    c.cache_mask = 0;
    if (cache_log2 > 0) {
        const size_t slots = (size_t)1 << cache_log2;
        m->store.emplace_back(slots * sizeof(kv::CacheEntry) + 128, 0);
        uintptr_t p = (reinterpret_cast<uintptr_t>(m->store.back().data()) + 127) & ~(uintptr_t)127;
        A.cache = reinterpret_cast<kv::CacheEntry*>(p);
        c.cache_mask = (uint32_t)(slots - 1);
    }
    return m;
}

static void emu_mcts_reset(EmuMcts* m, const uint64_t* start, uint64_t id_base) {
    for (int g = 0; g < m->G; g++) {
        kv::GameHdr h;
        memset(&h, 0, sizeof(h));
        h.game_id = id_base + (uint64_t)g;
        m->A.hdr[g] = h;
        for (int j = 0; j < m->cfg.inflight; j++) m->A.pend_node[(size_t)g * m->cfg.inflight + j] = -1;
        for (int i = 0; i < 16; i++) m->A.root_line[(size_t)g * 16 + i] = i >= 13 ? 0 : start[(size_t)g * 16 + i];
    }
}

static uint32_t g_emu_wave = 0;
static long g_emu_evals = 0, g_emu_late = 0;

static int g_emu_pipe = 0;   // 0: one group; 1..3: two groups, one of the interleavings kv_mcts.cu's events allow

// the view of the arrays a group's launches get (mirrors mcts_wave_group in kv_mcts.cu)
static kv::MctsArrays emu_view(EmuMcts* m, int q, int s0, uint32_t peer_wave) {
    kv::MctsArrays A = m->A;
    A.n_eval += q;
    A.n_late += q;
    A.eval_game += s0;
    A.eval_lines += (size_t)s0 * 16;
    A.eval_centry += s0;
    A.eval_hash += s0;
    A.late_game += s0;
    A.late_src += s0;
    A.slot_base = s0;
    A.peer_wave = peer_wave;
    return A;
}

struct EmuGroup {
    int g0 = 0, g1 = 0;
    kv::MctsArrays A;
    uint32_t wave = 0;
};

static void emu_select(EmuMcts* m, EmuGroup& G) {
    uint16_t mv[kv::MAX_MOVES];
    *G.A.n_eval = 0;
    *G.A.n_late = 0;
    for (int g = G.g0; g < G.g1; g++)
        kvemu::run_warp([&](int lane) { kv::mcts_select_warp(g_tables, lane, m->cfg, G.A, g, mv, G.wave); });
}
static void emu_eval(EmuMcts* m, EmuGroup& G) {
    float scratch[kv::MAX_MOVES];
    const int ne = (int)*G.A.n_eval;
    g_emu_evals += ne;
    for (int slot = 0; slot < ne; slot++)
        kvemu::run_warp([&](int lane) { kv::mcts_hash_eval_warp(lane, m->cfg, G.A, slot, scratch, G.wave); });
}
static void emu_late(EmuMcts* m, EmuGroup& G) {
    float scratch[kv::MAX_MOVES];
    const int nl = (int)*G.A.n_late;
    g_emu_late += nl;
    for (int li = 0; li < nl; li++)
        kvemu::run_warp([&](int lane) { kv::mcts_hash_late_warp(lane, m->cfg, G.A, li, scratch); });
    if (m->cfg.inflight > 1)
        for (int g = G.g0; g < G.g1; g++)
            kvemu::run_warp([&](int lane) { kv::mcts_backup_game_warp(lane, m->cfg, G.A, g); });
}

static void emu_mcts_wave(EmuMcts* m) {
    EmuGroup G;
    G.g1 = m->G;
    G.wave = ++g_emu_wave;
    G.A = emu_view(m, 0, 0, 0);
    emu_select(m, G);
    emu_eval(m, G);
    emu_late(m, G);
}

// n waves for both groups in one of the orders the device's stream/event graph admits (kv_mcts.cu mcts_run_waves):
//   1: A and B strictly one after the other          selA evA ltA | selB evB ltB
//   2: both selections first                          selA selB | evA ltA evB ltB      (B follows A's pending leaders)
//   3: A runs one selection ahead                     ... selB(w) evA(w) ltA(w) selA(w+1) evB(w) ltB(w)
//                                                     (A(w+1) follows B(w)'s pending leaders as well)
static void emu_mcts_waves_piped(EmuMcts* m, int n, int order) {
    const int K = m->cfg.inflight;
    const int gh = ((m->G / 2) + 7) & ~7;
    EmuGroup A, B;
    A.g0 = 0; A.g1 = gh;
    B.g0 = gh; B.g1 = m->G;
    uint32_t lastA = 0, lastB = 0;
    auto begin = [&](EmuGroup& X, int q, uint32_t peer) {
        X.wave = ++g_emu_wave;
        X.A = emu_view(m, q, X.g0 * K, peer);
    };
    if (order == 3) {
        begin(A, 0, 0);
        lastA = A.wave;
        emu_select(m, A);
    }
    for (int w = 0; w < n; w++) {
        if (order == 1) {
            begin(A, 0, lastB); lastA = A.wave;
            emu_select(m, A); emu_eval(m, A); emu_late(m, A);
            begin(B, 1, lastA); lastB = B.wave;
            emu_select(m, B); emu_eval(m, B); emu_late(m, B);
        } else if (order == 2) {
            begin(A, 0, lastB); lastA = A.wave;
            emu_select(m, A);
            begin(B, 1, lastA); lastB = B.wave;
            emu_select(m, B);
            emu_eval(m, A); emu_late(m, A);
            emu_eval(m, B); emu_late(m, B);
        } else {
            begin(B, 1, lastA); lastB = B.wave;
            emu_select(m, B);
            emu_eval(m, A); emu_late(m, A);
            if (w + 1 < n) {
                EmuGroup A2 = A;
                begin(A2, 0, lastB); lastA = A2.wave;
                emu_select(m, A2);
                emu_eval(m, B); emu_late(m, B);
                A = A2;
            } else {
                emu_eval(m, B); emu_late(m, B);
            }
        }
    }
}

// the driver loop of kv_mcts_run_move: waves until every live game has run its simulations
static void emu_mcts_move_waves(EmuMcts* m) {
    const bool piped = g_emu_pipe > 0 && m->G >= 16;
    const int K = m->cfg.inflight, S = m->cfg.sims;
    int n = K == 1 ? S : 1 + (S - 1 + K - 1) / K;
    for (;;) {
        bool left = false;
        for (int g = 0; g < m->G; g++) left |= !m->A.hdr[g].done && m->A.hdr[g].sims_done < m->cfg.sims;
        if (!left) break;
        if (piped) {
            emu_mcts_waves_piped(m, n, g_emu_pipe);
            n = 2;
        } else {
            emu_mcts_wave(m);
        }
    }
}

static void emu_mcts_finish(EmuMcts* m) {
    uint16_t mv[kv::MAX_MOVES];
    for (int g = 0; g < m->G; g++)
        kvemu::run_warp([&](int lane) { kv::mcts_finish_move_warp(g_tables, lane, m->cfg, m->A, g, mv); });
}

extern "C" {

__attribute__((visibility("default"))) void kvemu_set_rules(float resign_thr, int resign_min_plies, int root_mix) {
    g_emu_resign_thr = resign_thr;
    g_emu_resign_min = resign_min_plies;
    g_emu_root_mix = root_mix;
}
__attribute__((visibility("default"))) void kvemu_set_script(const uint16_t* moves, const float* vals, int stride) {
    g_emu_script_moves = moves;
    g_emu_script_vals = vals;
    g_emu_script_stride = stride;
}

// One search (sims waves) for G root positions; outputs the root edges of every game ([G][256]) and info [G][4]
__attribute__((visibility("default"))) void kvemu_mcts_search(int G, const uint64_t* start, uint64_t id_base, int ply0,
                                                               int sims, int edges_per_node, float c_puct,
                                                               float dir_alpha, float dir_eps, uint64_t seed,
                                                               uint16_t* moves, uint32_t* N, float* W, float* P,
                                                               int32_t* info, int inflight) {
    EmuMcts* m = emu_mcts_new(G, sims, edges_per_node, 1 << 20, 0, c_puct, dir_alpha, dir_eps, seed, 0, inflight);
    m->cfg.rec_cap = 1;
    emu_mcts_reset(m, start, id_base);
    for (int g = 0; g < G; g++) m->A.hdr[g].ply = ply0;
    emu_mcts_move_waves(m);
    for (int g = 0; g < G; g++) {
        const kv::GameHdr& h = m->A.hdr[g];
        const kv::NodeMeta nm = m->A.node_meta[(size_t)g * m->cfg.node_cap];
        const int n = (h.n_nodes && !(nm.ne_term & kv::NODE_TERM)) ? (nm.ne_term & 0xFFFF) : 0;
        info[4 * g + 0] = n;
        info[4 * g + 1] = h.n_nodes;
        info[4 * g + 2] = h.n_edges;
        info[4 * g + 3] = h.overflow;
        const size_t e0 = (size_t)g * m->cfg.edge_cap + (n ? nm.first_edge : 0);
        for (int k = 0; k < n; k++) {
            moves[256 * g + k] = m->A.eMv[e0 + k];
            N[256 * g + k] = m->A.eN[e0 + k];
            W[256 * g + k] = m->A.eW[e0 + k];
            P[256 * g + k] = m->A.eP[e0 + k];
        }
    }
    delete m;
}

// Whole games; out_moves [G][max_plies], out_plies [G], out_result [G]
__attribute__((visibility("default"))) void kvemu_selfplay(int G, const uint64_t* start, uint64_t id_base, int sims,
                                                            int edges_per_node, int max_plies, int temp_plies,
                                                            float c_puct, float dir_alpha, float dir_eps, uint64_t seed,
                                                            uint16_t* out_moves, int32_t* out_plies, int32_t* out_result,
                                                            int cache_log2, int64_t* out_counts2, int inflight,
                                                            int pipe_order) {
    EmuMcts* m = emu_mcts_new(G, sims, edges_per_node, max_plies, temp_plies, c_puct, dir_alpha, dir_eps, seed, cache_log2,
                              inflight);
    g_emu_pipe = pipe_order;
    g_emu_evals = g_emu_late = 0;
    emu_mcts_reset(m, start, id_base);
    for (int mvno = 0; mvno < max_plies; mvno++) {
        bool live = false;
        for (int g = 0; g < G; g++) live |= !m->A.hdr[g].done;
        if (!live) break;
        emu_mcts_move_waves(m);
        emu_mcts_finish(m);
    }
    for (int g = 0; g < G; g++) {
        out_plies[g] = m->A.hdr[g].ply;
        out_result[g] = (m->A.hdr[g].result & 0xFF) | (m->A.hdr[g].overflow << 8);   // low byte: result (two's complement), bits 8+: flags
        for (int p = 0; p < m->A.hdr[g].ply && p < max_plies; p++)
            out_moves[(size_t)g * max_plies + p] = m->A.rec_move[(size_t)g * m->cfg.rec_cap + p];
    }
    if (out_counts2) {
        out_counts2[0] = g_emu_evals;
        out_counts2[1] = g_emu_late;
    }
    g_emu_pipe = 0;
    delete m;
}

// task order of the whole-tower launch: out [total][3] = (layer, board tile, channel tile) of task t; returns total
__attribute__((visibility("default"))) int kvemu_tower_order(int M, int NT, int n_layers, int chunk_tiles, int32_t* out) {
    kvn::TowerOrder o;
    o.init(M, NT, n_layers, chunk_tiles);
    for (int t = 0; out && t < o.total; t++) {
        int l, m, n;
        o.decode(t, l, m, n);
        out[3 * t] = l; out[3 * t + 1] = m; out[3 * t + 2] = n;
    }
    return o.total;
}

__attribute__((visibility("default"))) void kvemu_tower_slice(int part, int num, int den, int M, int32_t* out2) {
    int off, cnt;
    kvn::tower_slice_tiles(part, num, den, M, off, cnt);
    out2[0] = off;
    out2[1] = cnt;
}

}  // extern "C"
