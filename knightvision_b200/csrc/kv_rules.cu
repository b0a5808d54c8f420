// kv_rules.cu — rules kernels (move generation, make-move, perft, encode) and their C-ABI entry points.
//
// The rules kernels give 8 lanes to a board (four boards per warp, kv_rules.cuh): the 128-byte board line is two
// coalesced 64 B requests per board (lane q keeps words q and q + 8), attack tables (5.5 KB) are staged once per CTA
// in shared memory, every board in flight has its own move buffer in shared memory.
// They are integer/bit kernels: HBM traffic is ~170 B per board, so they are bounded by issue rate,
// not bandwidth (DESIGN.md §kernels gives both figures).
#include <cstring>
#include <vector>

#include "kv_internal.h"
#include "kv_rules.cuh"
#include "kv_tables_dev.cuh"

namespace kv {


constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
#ifndef KV_RULES_MIN_CTAS
#define KV_RULES_MIN_CTAS 3                  // resident CTAs per SM the register budget is sized for (tuning knob)
#endif
#ifndef KV_RULES_LANES
#define KV_RULES_LANES 8                     // lanes per board in the rules kernels: 8 = four boards per warp (16: two)
#endif
constexpr int kW = KV_RULES_LANES;
constexpr int kBoardsPerWarp = 32 / kW;
constexpr int kGroupsPerCta = kWarpsPerCta * kBoardsPerWarp;

struct __align__(16) RulesSmem {
    Tables tab;
    uint16_t mv[kGroupsPerCta][MAX_MOVES];   // one move buffer per board in flight
};

__device__ __forceinline__ void stage_tables(RulesSmem& sm) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(&g_tables);
    uint64_t* dst = reinterpret_cast<uint64_t*>(&sm.tab);
    for (int i = threadIdx.x; i < kTableWords; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// contiguous slice of [0,n) for global warp gw of nw
__device__ __forceinline__ void warp_slice(int n, int gw, int nw, int& lo, int& hi) {
    const int per = (n + nw - 1) / nw;
    lo = gw * per;
    hi = lo + per < n ? lo + per : n;
}

using RLine = Line<kW>;

// the words of the board line lane q of a group keeps (q, q + WL, ...); zeros when the group has no board
__device__ __forceinline__ RLine ld_line(const uint64_t* line, int q, bool valid) {
    RLine L;
#pragma unroll
    for (int j = 0; j < RLine::NW; j++) {
        const int idx = q + j * RLine::WL;
        L.w[j] = (valid && idx < LINE_WORDS) ? __ldg(line + idx) : 0ull;
    }
    return L;
}

// (2 CTAs per SM: with 128 registers the ordered-list path does not spill — measured 610 against 490 M boards/s at 3)
__global__ void __launch_bounds__(kThreads, 2) movegen_kernel(uint64_t* __restrict__ lines, int n,
                                                           uint16_t* __restrict__ moves, int stride,
                                                           int32_t* __restrict__ counts, int32_t* __restrict__ flags) {
    __shared__ RulesSmem sm;
    stage_tables(sm);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = lane & (kW - 1), grp = lane / kW;
    int lo, hi;
    warp_slice(n, blockIdx.x * kWarpsPerCta + wid, gridDim.x * kWarpsPerCta, lo, hi);
    uint16_t* mv = sm.mv[wid * kBoardsPerWarp + grp];
    for (int i0 = lo; i0 < hi; i0 += kBoardsPerWarp) {
        const int i = i0 + grp;
        const bool valid = i < hi;
        uint64_t* line = lines + (size_t)(valid ? i : i0) * LINE_WORDS;
        RLine L = ld_line(line, q, valid);
        const GenOut g = movegen_sub<kW>(sm.tab, lane, L, mv);
        if (valid) {
            if (g.flags & RF_STATE_MUTATED) {
#pragma unroll
                for (int j = 0; j < RLine::NW; j++)
                    if (q + j * RLine::WL < 12) line[q + j * RLine::WL] = L.w[j];
            }
            int cnt = g.n < stride ? g.n : stride;
            if (cnt > MAX_MOVES) cnt = MAX_MOVES;
            // packed u32 stores
            uint32_t* dst = reinterpret_cast<uint32_t*>(moves + (size_t)i * stride);
            for (int k = q; 2 * k < cnt; k += kW) {
                uint32_t lo16 = mv[2 * k], hi16 = (2 * k + 1 < cnt) ? mv[2 * k + 1] : 0u;
                dst[k] = lo16 | (hi16 << 16);
            }
            if (q == 0) {
                counts[i] = g.n;
                flags[i] = g.flags | ((g.n > stride) ? RF_OVERFLOW : 0);
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kThreads, KV_RULES_MIN_CTAS) make_moves_kernel(uint64_t* __restrict__ lines, int n,
                                                              const uint16_t* __restrict__ mvs) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = lane & (kW - 1), grp = lane / kW;
    int lo, hi;
    warp_slice(n, blockIdx.x * kWarpsPerCta + wid, gridDim.x * kWarpsPerCta, lo, hi);
    for (int i0 = lo; i0 < hi; i0 += kBoardsPerWarp) {
        const int i = i0 + grp;
        const int m = i < hi ? mvs[i] : 0xFFFF;
        const bool valid = m != 0xFFFF;
        uint64_t* line = lines + (size_t)(i < hi ? i : i0) * LINE_WORDS;
        RLine L = ld_line(line, q, valid);
        make_move_sub<kW>(lane, L, valid ? m : 0, T_Q);
        if (valid) {
#pragma unroll
            for (int j = 0; j < RLine::NW; j++)
                if (q + j * RLine::WL < 13) line[q + j * RLine::WL] = L.w[j];
        }
    }
}

__global__ void __launch_bounds__(kThreads, KV_RULES_MIN_CTAS) attacked_kernel(const uint64_t* __restrict__ lines, int n,
                                                            uint64_t* __restrict__ masks) {
    __shared__ RulesSmem sm;
    stage_tables(sm);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int lo, hi;
    warp_slice(n, blockIdx.x * kWarpsPerCta + wid, gridDim.x * kWarpsPerCta, lo, hi);
    for (int i = lo; i < hi; i++) {
        const uint64_t w = lane < LINE_WORDS ? __ldg(lines + (size_t)i * LINE_WORDS + lane) : 0ull;
        const uint64_t m = attacked_mask_warp(sm.tab, lane, w);
        if (lane == 0) masks[i] = m;
    }
}

// ---- perft: one frontier level per launch (body: perft_visit_sub, kv_rules.cuh) ---------------------------
template <bool LEAF, bool DIGEST>
__global__ void __launch_bounds__(kThreads, KV_RULES_MIN_CTAS) perft_level_kernel(const uint64_t* __restrict__ cur, int m,
                                                               uint64_t* __restrict__ next,
                                                               uint32_t* __restrict__ next_count,
                                                               uint64_t* __restrict__ out) {
    __shared__ RulesSmem sm;
    stage_tables(sm);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = lane & (kW - 1), grp = lane / kW;
    int lo, hi;
    warp_slice(m, blockIdx.x * kWarpsPerCta + wid, gridDim.x * kWarpsPerCta, lo, hi);
    uint64_t accv = 0;
    int acc_root = -1;
    uint16_t* mv = sm.mv[wid * kBoardsPerWarp + grp];
    // the next group of lines is requested before the current one is visited
    RLine Ln = ld_line(cur + (size_t)(lo + grp < hi ? lo + grp : lo) * LINE_WORDS, q, lo + grp < hi);
    for (int i0 = lo; i0 < hi; i0 += kBoardsPerWarp) {
        const bool valid = i0 + grp < hi;
        const RLine L = Ln;
        const int in = i0 + kBoardsPerWarp + grp;
        Ln = ld_line(cur + (size_t)(in < hi ? in : lo) * LINE_WORDS, q, in < hi);
        perft_visit_sub<kW, LEAF, DIGEST>(sm.tab, lane, L, valid, mv, accv, acc_root, next, next_count, out);
    }
    perft_acc_flush(accv, acc_root, out, q);
}

__global__ void perft_seed_kernel(const uint64_t* __restrict__ roots, int n, uint64_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * LINE_WORDS) return;
    const int b = i / LINE_WORDS, wd = i % LINE_WORDS;
    uint64_t v = roots[i];
    if (wd == 13) v = (uint64_t)(uint32_t)b;
    if (wd == 14 || wd == 15) v = 0;
    dst[i] = v;
}

// encode_board (ai/ai.py:17-41): one thread per (board, plane, row) writes 8 floats (two float4 stores)
__global__ void encode_kernel(const uint64_t* __restrict__ lines, int n, float* __restrict__ planes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * 96) return;
    const size_t b = i / 96;
    const int pr = (int)(i % 96), pl = pr >> 3, row = pr & 7;
    const uint32_t bits = (uint32_t)((__ldg(lines + b * LINE_WORDS + pl) >> (row * 8)) & 0xFF);
    float4 a, c;
    a.x = (bits >> 0) & 1; a.y = (bits >> 1) & 1; a.z = (bits >> 2) & 1; a.w = (bits >> 3) & 1;
    c.x = (bits >> 4) & 1; c.y = (bits >> 5) & 1; c.z = (bits >> 6) & 1; c.w = (bits >> 7) & 1;
    float4* dst = reinterpret_cast<float4*>(planes + i * 8);
    dst[0] = a;
    dst[1] = c;
}

static int grid_for(kv_ctx* ctx, int n) {
    int g = (n + kGroupsPerCta - 1) / kGroupsPerCta;
    const int cap = ctx->sm_count * 8;   // 8 CTAs x 8 warps = 64 resident warps per SM
    if (g > cap) g = cap;
    return g < 1 ? 1 : g;
}

}  // namespace kv

using namespace kv;

extern "C" {

int kv_movegen(kv_ctx* ctx, uint64_t* d_lines, int n, uint16_t* d_moves, int stride, int32_t* d_counts,
               int32_t* d_flags, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    if (stride < 2 || (stride & 1)) return kv_fail_msg(ctx, "kv_movegen: stride must be even and >= 2");
    KvTimed t_(ctx, KVK_MOVEGEN, (cudaStream_t)stream);
    movegen_kernel<<<grid_for(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(d_lines, n, d_moves, stride, d_counts, d_flags);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_make_moves(kv_ctx* ctx, uint64_t* d_lines, int n, const uint16_t* d_moves, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    KvTimed t_(ctx, KVK_MAKE_MOVES, (cudaStream_t)stream);
    make_moves_kernel<<<grid_for(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(d_lines, n, d_moves);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_attacked(kv_ctx* ctx, const uint64_t* d_lines, int n, uint64_t* d_masks, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    attacked_kernel<<<grid_for(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(d_lines, n, d_masks);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_encode(kv_ctx* ctx, const uint64_t* d_lines, int n, float* d_planes, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    const size_t items = (size_t)n * 96;
    KvTimed t_(ctx, KVK_ENCODE, (cudaStream_t)stream);
    encode_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lines, n, d_planes);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

// Depth-first over chunks of at most `chunk` boards per launch, breadth-first inside a chunk.
static int perft_rec(kv_ctx* ctx, const uint64_t* cur, int m, int remaining, int level, uint64_t* d_out, int chunk,
                     cudaStream_t st, bool digest) {
    for (int lo = 0; lo < m; lo += chunk) {
        const int c = (m - lo) < chunk ? (m - lo) : chunk;
        const uint64_t* src = cur + (size_t)lo * LINE_WORDS;
        if (remaining == 1) {
            {
                KvTimed t_(ctx, KVK_PERFT_LEAF, st);
                if (digest) perft_level_kernel<true, true><<<grid_for(ctx, c), kThreads, 0, st>>>(src, c, nullptr, nullptr, d_out);
                else perft_level_kernel<true, false><<<grid_for(ctx, c), kThreads, 0, st>>>(src, c, nullptr, nullptr, d_out);
            }
            KV_LAUNCH_CHECK(ctx);
        } else {
            KV_CUDA(ctx, cudaMemsetAsync(ctx->perft_counter + level, 0, sizeof(uint32_t), st));
            {
                KvTimed t_(ctx, KVK_PERFT_EXPAND, st);
                if (digest)
                    perft_level_kernel<false, true><<<grid_for(ctx, c), kThreads, 0, st>>>(
                        src, c, ctx->perft_buf[level], ctx->perft_counter + level, d_out);
                else
                    perft_level_kernel<false, false><<<grid_for(ctx, c), kThreads, 0, st>>>(
                        src, c, ctx->perft_buf[level], ctx->perft_counter + level, d_out);
            }
            KV_LAUNCH_CHECK(ctx);
            uint32_t cnt = 0;
            KV_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->perft_counter + level, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            KV_CUDA(ctx, cudaStreamSynchronize(st));
            if (cnt) {
                int rc = perft_rec(ctx, ctx->perft_buf[level], (int)cnt, remaining - 1, level + 1, d_out, chunk, st, digest);
                if (rc) return rc;
            }
        }
    }
    return 0;
}

int kv_perft(kv_ctx* ctx, const uint64_t* d_roots, int n, int depth, uint64_t* d_out, int chunk, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    // chunk < 0: counts only (nodes, categories, movegen calls) — the order digest stays 0 and leaves are bulk-counted
    // from the destination sets without laying the move lists out
    const bool digest = chunk >= 0;
    if (chunk < 0) chunk = -chunk;
    if (depth < 1 || depth > 8) return kv_fail_msg(ctx, "kv_perft: depth must be in 1..8");
    if (chunk <= 0) chunk = 65536;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = (size_t)chunk * MAX_MOVES;   // a chunk can never produce more children than this
    if (ctx->perft_cap < cap) {
        for (int l = 1; l < 8; l++) {
            if (ctx->perft_buf[l]) cudaFree(ctx->perft_buf[l]);
            ctx->perft_buf[l] = nullptr;
        }
        ctx->perft_cap = cap;
    }
    for (int l = 1; l < depth; l++)
        if (!ctx->perft_buf[l])
            KV_CUDA(ctx, cudaMalloc(&ctx->perft_buf[l], ctx->perft_cap * LINE_WORDS * sizeof(uint64_t)));
    if (ctx->perft_roots_cap < (size_t)n) {
        if (ctx->perft_buf[0]) cudaFree(ctx->perft_buf[0]);
        ctx->perft_buf[0] = nullptr;
        KV_CUDA(ctx, cudaMalloc(&ctx->perft_buf[0], (size_t)n * LINE_WORDS * sizeof(uint64_t)));
        ctx->perft_roots_cap = (size_t)n;
    }
    if (!ctx->perft_counter) KV_CUDA(ctx, cudaMalloc(&ctx->perft_counter, 16 * sizeof(uint32_t)));
    KV_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * 8 * sizeof(uint64_t), st));
    perft_seed_kernel<<<(n * LINE_WORDS + 255) / 256, 256, 0, st>>>(d_roots, n, ctx->perft_buf[0]);
    KV_LAUNCH_CHECK(ctx);
    return perft_rec(ctx, ctx->perft_buf[0], n, depth, 1, d_out, chunk, st, digest);
}

}  // extern "C"
