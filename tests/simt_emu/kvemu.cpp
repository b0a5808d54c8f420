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

static const kv::Tables g_tables = kv::make_tables();

extern "C" {

__attribute__((visibility("default"))) void kvemu_movegen(uint64_t* lines, int n, uint16_t* moves, int stride,
                                                           int32_t* counts, int32_t* flags) {
    for (int i = 0; i < n; i++) {
        uint16_t mv[kv::MAX_MOVES];
        memset(mv, 0, sizeof(mv));
        uint64_t* line = lines + 16 * (size_t)i;
        kv::GenOut out[32];
        uint64_t neww[32];
        kvemu::run_warp([&](int lane) {
            uint64_t w = lane < 16 ? line[lane] : 0;
            out[lane] = kv::movegen_warp(g_tables, lane, w, mv);
            neww[lane] = w;
        });
        for (int l = 1; l < 32; l++)
            if (out[l].n != out[0].n || out[l].flags != out[0].flags) {
                fprintf(stderr, "kvemu: non-uniform movegen result\n");
                abort();
            }
        if (out[0].flags & kv::RF_STATE_MUTATED)
            for (int l = 0; l < 16; l++) line[l] = neww[l];
        counts[i] = out[0].n;
        flags[i] = out[0].flags;
        for (int k = 0; k < out[0].n && k < stride && k < kv::MAX_MOVES; k++) moves[(size_t)i * stride + k] = mv[k];
    }
}

__attribute__((visibility("default"))) void kvemu_make_moves(uint64_t* lines, int n, const uint16_t* mv) {
    for (int i = 0; i < n; i++) {
        uint64_t* line = lines + 16 * (size_t)i;
        uint64_t neww[32];
        kvemu::run_warp([&](int lane) {
            uint64_t w = lane < 16 ? line[lane] : 0;
            neww[lane] = kv::make_move_warp(lane, w, mv[i], kv::T_Q);
        });
        for (int l = 0; l < 16; l++) line[l] = neww[l];
    }
}

// kv_perft's level loop over host memory (same chunked depth-first / breadth-first-in-chunk order)
static void emu_perft_rec(std::vector<uint64_t>& cur, int remaining, uint64_t* out) {
    const int m = (int)(cur.size() / 16);
    uint16_t mv[kv::MAX_MOVES];
    if (remaining == 1) {
        for (int i = 0; i < m; i++) {
            uint64_t acc[32] = {0};
            kvemu::run_warp([&](int lane) {
                int acc_root = -1;
                uint64_t w = lane < 16 ? cur[16 * (size_t)i + lane] : 0;
                kv::perft_visit_warp<true>(g_tables, lane, w, mv, acc[lane], acc_root, nullptr, nullptr, out);
                kv::perft_acc_flush(acc[lane], acc_root, out, lane);
            });
        }
        return;
    }
    std::vector<uint64_t> next((size_t)m * kv::MAX_MOVES * 16);
    uint32_t cnt = 0;
    for (int i = 0; i < m; i++) {
        uint64_t acc[32] = {0};
        kvemu::run_warp([&](int lane) {
            int acc_root = -1;
            uint64_t w = lane < 16 ? cur[16 * (size_t)i + lane] : 0;
            kv::perft_visit_warp<false>(g_tables, lane, w, mv, acc[lane], acc_root, next.data(), &cnt, out);
            kv::perft_acc_flush(acc[lane], acc_root, out, lane);
        });
    }
    next.resize((size_t)cnt * 16);
    if (cnt) emu_perft_rec(next, remaining - 1, out);
}

__attribute__((visibility("default"))) void kvemu_perft(const uint64_t* roots, int n, int depth, uint64_t* out) {
    memset(out, 0, (size_t)n * 8 * sizeof(uint64_t));
    std::vector<uint64_t> cur(roots, roots + (size_t)n * 16);
    for (int i = 0; i < n; i++) {
        cur[16 * (size_t)i + 13] = (uint32_t)i;
        cur[16 * (size_t)i + 14] = 0;
        cur[16 * (size_t)i + 15] = 0;
    }
    emu_perft_rec(cur, depth, out);
}

}  // extern "C"
