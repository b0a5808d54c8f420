// kv_mcts.cu — self-play engine: batched PUCT search kernels (kv_mcts.cuh), the search-mode network evaluator
// (legal-move logits + masked softmax fused with the heads), game records, and their C-ABI entry points.
//
// One search wave = one simulation for every live game:
//   select_kernel      warp per game: descend / make-move / movegen / queue leaf       (integer + fp32 PUCT)
//   network            stem + tcgen05 tower over the queued leaves (count stays on the device)      (kv_net.cu)
//   eval_net_kernel    CTA per leaf: heads, logits of the LEGAL moves only, softmax priors, root noise, backup
// cfg.sims waves, then finish_move_kernel picks / records / plays the move for every game.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kv_heads.cuh"
#include "kv_internal.h"
#include "kv_mcts.cuh"
#include "kv_net.h"
#include "kv_tables_dev.cuh"

namespace kv {


constexpr int kMW = 8;   // warps (games) per CTA
constexpr int kMWP = 2;  // ... of the selection kernel in the pipelined search: 64-thread CTAs (8 K registers, 7.5 KB of
                         // shared memory) fit next to a resident tower CTA
template <int MW>
struct __align__(16) MctsSmemT {
    Tables tab;
    uint16_t mv[MW][MAX_MOVES];
};
using MctsSmem = MctsSmemT<kMW>;

__device__ __forceinline__ void stage_tables_m(Tables& dstT) {
    const uint64_t* src = reinterpret_cast<const uint64_t*>(&g_tables);
    uint64_t* dst = reinterpret_cast<uint64_t*>(&dstT);
    for (int i = threadIdx.x; i < kTableWords; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// games [g0, g1): the whole context, or one of the two groups of the pipelined search
template <int MW>
__global__ void __launch_bounds__(MW * 32, 16 / MW) mcts_select_kernel(MctsCfg cfg, MctsArrays A, int g0, int g1, uint32_t wave) {
    __shared__ MctsSmemT<MW> sm;
    stage_tables_m(sm.tab);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // grid-stride over the games: the pipelined search launches only as many CTAs as fit NEXT TO the resident tower
    // CTAs (so they never keep the next tower layer waiting for an SM); otherwise one game per warp, one pass
    for (int g = g0 + blockIdx.x * MW + wid; g < g1; g += gridDim.x * MW)
        mcts_select_warp(sm.tab, lane, cfg, A, g, sm.mv[wid], wave);
}

__global__ void __launch_bounds__(kMW * 32) mcts_hash_eval_kernel(MctsCfg cfg, MctsArrays A, uint32_t wave) {
    __shared__ float scratch[kMW][MAX_MOVES];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int slot = blockIdx.x * kMW + wid;
    if (slot >= (int)*A.n_eval) return;
    mcts_hash_eval_warp(lane, cfg, A, slot, scratch[wid], wave);
}

__global__ void __launch_bounds__(kMW * 32) mcts_hash_late_kernel(MctsCfg cfg, MctsArrays A) {
    __shared__ float scratch[kMW][MAX_MOVES];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int li = blockIdx.x * kMW + wid;
    if (li >= (int)*A.n_late) return;
    mcts_hash_late_warp(lane, cfg, A, li, scratch[wid]);
}

// K > 1 only: warp per game, backs the wave's pending simulations up in slot order (after every expansion)
__global__ void __launch_bounds__(kMW * 32) mcts_backup_kernel(MctsCfg cfg, MctsArrays A, int g0, int g1) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int g = g0 + blockIdx.x * kMW + wid; g < g1; g += gridDim.x * kMW) mcts_backup_game_warp(lane, cfg, A, g);
}

__global__ void __launch_bounds__(kMW * 32) mcts_finish_move_kernel(MctsCfg cfg, MctsArrays A, int G) {
    __shared__ MctsSmem sm;
    stage_tables_m(sm.tab);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = blockIdx.x * kMW + wid;
    if (g >= G) return;
    mcts_finish_move_warp(sm.tab, lane, cfg, A, g, sm.mv[wid]);
}

struct HeadW {
    const float *wh, *bh, *wfc, *bfc, *w1, *b1, *w2, *b2;
    int C;
};

// logits of the pending leaf's legal moves from the 128 policy features (policy_fc rows of the legal indices only).
// Eight consecutive lanes share a move: each takes 16 of the 128 features (four float4 of the weight row), then a
// fixed-order butterfly over the 8 lanes — the same code serves tower evaluations and cache hits, so a hit
// reproduces a miss bit for bit.
__device__ __forceinline__ void legal_logits(const MctsCfg& cfg, const MctsArrays& A, int gs, const HeadW& H, const float* hp,
                                             float* logits) {
    const int g = gs / cfg.inflight;
    const NodeMeta m = A.node_meta[(size_t)g * cfg.node_cap + A.pend_node[gs]];
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = (size_t)g * cfg.edge_cap + m.first_edge;
    const int part = threadIdx.x & 7, per = blockDim.x >> 3;
    for (int k0 = 0; k0 < n; k0 += per) {   // uniform trip count: the shuffles below need whole warps
        const int k = k0 + (int)(threadIdx.x >> 3);
        float a = 0.f;
        int idx = 0;
        if (k < n) {
            idx = move_index(A.eMv[e0 + k]);
            const float4* wr = reinterpret_cast<const float4*>(H.wfc + (size_t)idx * 128 + part * 16);
            const float* f = hp + part * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 w = __ldg(wr + i);
                a += w.x * f[4 * i] + w.y * f[4 * i + 1] + w.z * f[4 * i + 2] + w.w * f[4 * i + 3];
            }
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (k < n && part == 0) logits[k] = a + __ldg(H.bfc + idx);
    }
}

// Sum of the 4096 Gamma variates of the reference-rule root noise (key = ply * 4096 + index) with every thread of the
// CTA, in an order that does not depend on the CTA size (the tower path runs 256 threads, the cache-hit path 128, and a
// hit must reproduce a miss bit for bit): 256 columns, column c sums indices c, c + 256, ... in order; a butterfly over
// each group of 32 columns; the eight group sums added in order.  All threads return the sum.  `red`: 8 floats.
__device__ __forceinline__ float root_noise_sum_cta(const MctsCfg& cfg, const MctsArrays& A, int gs, float* red) {
    const GameHdr* h = &A.hdr[gs / cfg.inflight];
    const uint64_t key = (uint64_t)h->ply * POLICY_N;
    __syncthreads();                       // `red` may still be read by its previous user
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {      // uniform trip count per warp (blockDim is 128 or 256)
        float part = 0.f;
        for (int i = c; i < POLICY_N; i += 256)
            part = part + kvd_gamma_small(cfg.dir_alpha, cfg.seed, h->game_id, key + (uint64_t)i);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) part = part + __shfl_xor_sync(0xffffffffu, part, d);
        if ((c & 31) == 0) red[c >> 5] = part;
    }
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; w++) tot = tot + red[w];
    return tot;
}

// CTA per queued leaf: heads on the tower output, logits of the legal moves, then the warp-level expand/backup.
__global__ void __launch_bounds__(256) mcts_eval_net_kernel(MctsCfg cfg, MctsArrays A, const __nv_bfloat16* __restrict__ act,
                                                            HeadW H, uint32_t wave) {
    __shared__ float hp[128], hv[64], red[8], logits[MAX_MOVES];
    __shared__ __align__(16) float swh[3 * 512];
    const int n_eval = (int)*A.n_eval;
    for (int slot = blockIdx.x; slot < n_eval; slot += gridDim.x) {   // grid-stride: see mcts_select_kernel
        const int gs = A.eval_game[slot];
        kvn::head_features(act + (size_t)slot * 64 * H.C, H.C, H.wh, H.bh, hp, hv, swh);
        __syncthreads();
        const float v_white = kvn::value_mlp(hv, H.w1, H.b1, H.w2, H.b2, red);
        legal_logits(cfg, A, gs, H, hp, logits);
        __syncthreads();
        if (threadIdx.x < 32) {
            mcts_expand_warp((int)threadIdx.x, cfg, A, gs, logits, v_white);
            if (cfg.cache_mask) cache_fill_warp((int)threadIdx.x, cfg, A, wave, slot, hp, v_white);
        }
        __syncthreads();   // hp / hv / red / logits are reused by the next leaf
    }
}

// ---- the same evaluation as two kernels (opt-in: kv_mcts_set_eval_split(1) / KV_MCTS_EVAL_SPLIT=1) ----------------------
// ncu at bench state (profiles/r02_ncu_mcts_kernels_bench_state.txt) shows the one-CTA-per-leaf kernel moving 257 MB of
// last-layer activations at 1.6 TB/s and 481 MB from L2 to L1, most of it the 128 KB value-MLP matrix re-read by every
// leaf's CTA.  Split: (1) a streaming kernel that turns a leaf's 64 x C activations into its 128 policy + 64 value
// features (head_buf), (2) a kernel that finishes kEB leaves per CTA — the value MLP for all of them from one pass over
// w1, then one warp per leaf for the legal logits, priors and backup.  Same expressions in the same order as the kernel
// above: bit-identical results (tests/test_gpu_mcts.py::test_split_evaluator_matches_single_kernel).
// Measured at bench state (4 096 games x 800 sims, tools/bench_eval_split.py): expand kernels 184 -> 165 ms per move, and
// the tower 5 968 -> 5 990 ms — the step as a whole does not move (515.0 k vs 514.9 k sims/s): under the power cap the
// time the tree kernels give back is taken by a lower tower clock.  Hence opt-in.
constexpr int kEB = 8;      // leaves per CTA of the second kernel
constexpr int HEADF = 192;  // 128 policy features + 64 value features per leaf

__global__ void __launch_bounds__(256) mcts_head_kernel(MctsArrays A, const __nv_bfloat16* __restrict__ act, HeadW H,
                                                        float* __restrict__ head_buf) {
    __shared__ float hp[128], hv[64];
    __shared__ __align__(16) float swh[3 * 512];
    const int n_eval = (int)*A.n_eval;
    for (int slot = blockIdx.x; slot < n_eval; slot += gridDim.x) {
        kvn::head_features(act + (size_t)slot * 64 * H.C, H.C, H.wh, H.bh, hp, hv, swh);
        __syncthreads();
        if (threadIdx.x < HEADF)
            head_buf[(size_t)slot * HEADF + threadIdx.x] = threadIdx.x < 128 ? hp[threadIdx.x] : hv[threadIdx.x - 128];
        __syncthreads();
    }
}

// legal-move logits of one leaf by ONE warp (same per-move arithmetic as legal_logits: 8 lanes x 16 features, butterfly)
__device__ __forceinline__ void legal_logits_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int gs, const HeadW& H,
                                                  const float* hp, float* logits) {
    const int g = gs / cfg.inflight;
    const NodeMeta m = A.node_meta[(size_t)g * cfg.node_cap + A.pend_node[gs]];
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = (size_t)g * cfg.edge_cap + m.first_edge;
    const int part = lane & 7;
    for (int k0 = 0; k0 < n; k0 += 4) {   // uniform trip count: the shuffles below need the whole warp
        const int k = k0 + (lane >> 3);
        float a = 0.f;
        int idx = 0;
        if (k < n) {
            idx = move_index(A.eMv[e0 + k]);
            const float4* wr = reinterpret_cast<const float4*>(H.wfc + (size_t)idx * 128 + part * 16);
            const float* f = hp + part * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 w = __ldg(wr + i);
                a += w.x * f[4 * i] + w.y * f[4 * i + 1] + w.z * f[4 * i + 2] + w.w * f[4 * i + 3];
            }
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (k < n && part == 0) logits[k] = a + __ldg(H.bfc + idx);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) mcts_eval_batch_kernel(MctsCfg cfg, MctsArrays A, const float* __restrict__ head_buf,
                                                              HeadW H, uint32_t wave) {
    __shared__ float hf[kEB][HEADF], logits[kEB][MAX_MOVES], red[kEB][8], vw[kEB];
    const int n_eval = (int)*A.n_eval;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s0 = blockIdx.x * kEB; s0 < n_eval; s0 += gridDim.x * kEB) {
        const int nb = n_eval - s0 < kEB ? n_eval - s0 : kEB;
        for (int i = threadIdx.x; i < kEB * HEADF; i += blockDim.x) {
            const int l = i / HEADF;
            hf[l][i - l * HEADF] = l < nb ? head_buf[(size_t)(s0 + l) * HEADF + (i - l * HEADF)] : 0.f;
        }
        __syncthreads();
        // value_fc1 64 -> 512 + relu, value_fc2 512 -> 1 (kv_heads.cuh value_mlp, same association), all kEB leaves per
        // pass over the transposed w1: thread j owns units j and j + 256
        float part[kEB];
#pragma unroll
        for (int l = 0; l < kEB; l++) part[l] = 0.f;
        for (int j = threadIdx.x; j < 512; j += blockDim.x) {
            float a[kEB];
            const float bj = __ldg(H.b1 + j);
#pragma unroll
            for (int l = 0; l < kEB; l++) a[l] = bj;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float wx = __ldg(H.w1 + (4 * i + 0) * 512 + j), wy = __ldg(H.w1 + (4 * i + 1) * 512 + j);
                const float wz = __ldg(H.w1 + (4 * i + 2) * 512 + j), ww = __ldg(H.w1 + (4 * i + 3) * 512 + j);
#pragma unroll
                for (int l = 0; l < kEB; l++) {
                    const float* hv = hf[l] + 128;
                    a[l] += wx * hv[4 * i] + wy * hv[4 * i + 1] + wz * hv[4 * i + 2] + ww * hv[4 * i + 3];
                }
            }
            const float w2j = __ldg(H.w2 + j);
#pragma unroll
            for (int l = 0; l < kEB; l++) part[l] += fmaxf(a[l], 0.f) * w2j;
        }
#pragma unroll
        for (int l = 0; l < kEB; l++) {
            float p = part[l];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(0xffffffffu, p, m);
            if (lane == 0) red[l][wid] = p;
        }
        __syncthreads();
        if (threadIdx.x < kEB) {
            float tot = 0.f;
            for (int w = 0; w < 8; w++) tot += red[threadIdx.x][w];
            vw[threadIdx.x] = tanhf(tot + __ldg(H.b2));
        }
        __syncthreads();
        if (wid < nb) {   // warp l finishes leaf l
            const int slot = s0 + wid, gs = A.eval_game[slot];
            legal_logits_warp(lane, cfg, A, gs, H, hf[wid], logits[wid]);
            mcts_expand_warp(lane, cfg, A, gs, logits[wid], vw[wid]);
            if (cfg.cache_mask) cache_fill_warp(lane, cfg, A, wave, slot, hf[wid], vw[wid]);
        }
        __syncthreads();
    }
}

// CTA (128 threads) per late entry: features from the cache (already copied per game) or from this wave's leader
__global__ void __launch_bounds__(128) mcts_late_net_kernel(MctsCfg cfg, MctsArrays A, HeadW H) {
    __shared__ float hp[FEAT], logits[MAX_MOVES];
    const int n_late = (int)*A.n_late;
    for (int li = blockIdx.x; li < n_late; li += gridDim.x) {
        const int gs = A.late_game[li], src = A.late_src[li];
        const float* f = src < 0 ? A.feat_game + (size_t)gs * FEAT : A.feat_slot + (size_t)src * FEAT;
        for (int i = threadIdx.x; i < FEAT; i += blockDim.x) hp[i] = f[i];
        __syncthreads();
        legal_logits(cfg, A, gs, H, hp, logits);
        __syncthreads();
        if (threadIdx.x < 32) mcts_expand_warp((int)threadIdx.x, cfg, A, gs, logits, hp[128], true);
        __syncthreads();
    }
}

// ---- cfg.root_mix (the reference's prior rule, scripts/self_play.py:150-167): the root's softmax runs over ALL 4096
// logits and its Dirichlet noise over all 4096 indices.  These variants of the two evaluator kernels take kRM leaves per
// CTA pass, so that the 2 MB policy_fc matrix is read once per kRM leaves, and form
//   (mx, z)  softmax statistics over the 4096 logits: the logits of the kRM leaves go to shared memory (64 KB, dynamic),
//            then two warps per leaf form the maximum and the sum of exponentials (lane-strided, butterfly, halves in order)
//   gsum     the 4096-term Gamma sum (root_noise_sum_cta)
// Both kernels run 256 threads and the per-leaf arithmetic does not depend on which leaves share a pass, so a root served
// by the evaluation cache gets the very bits of a root that went through the tower.
constexpr int kRM = 4;
constexpr int kRootMixDyn = kRM * POLICY_N * (int)sizeof(float);   // dynamic shared memory of the two kernels (64 KB)
struct RootMixSmem {
    float hp[kRM][FEAT];
    float hv[64];
    float logits[kRM][MAX_MOVES];
    float pm[kRM][2], pz[kRM][2];
    float red[8];
    float vw[kRM], gsum[kRM], mx[kRM], z[kRM];
    int gs[kRM];
};

// lall: kRM x 4096 floats of dynamic shared memory (64 KB)
__device__ __forceinline__ void rootmix_stats(const MctsCfg& cfg, const MctsArrays& A, const HeadW& H, RootMixSmem& sm, int nb,
                                              float* lall) {
    const int part = threadIdx.x & 7, own = threadIdx.x >> 3;   // 8 lanes per row, 32 rows per pass (256 threads)
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // all 4096 logits of the kRM leaves: every policy_fc row is loaded once and used kRM times (fused multiply-adds: these
    // values only feed the softmax statistics, which no CPU code has to reproduce)
    // every CTA starts at its own row block (the logits are stored by index, so the order does not matter): otherwise
    // the ~300 resident CTAs walk the same 2 MB matrix in lock step and queue up on the same L2 lines; the next block's
    // weights are requested before the current block is used
    // the lane's 16 features of every leaf live in registers for the whole pass (read from shared memory inside the loop
    // they cost a 4-way bank conflict per load: eight 64-byte-strided addresses per warp fall into two bank groups)
    float f[kRM][16];
#pragma unroll
    for (int r = 0; r < kRM; r++)
#pragma unroll
        for (int j = 0; j < 16; j++) f[r][j] = sm.hp[r][part * 16 + j];
    const int start = (int)((blockIdx.x * 416u) & (POLICY_N - 1)) & ~31;
    float4 wn[4];
    {
        const float4* wr = reinterpret_cast<const float4*>(H.wfc + (size_t)(start + own) * 128 + part * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) wn[i] = __ldg(wr + i);
    }
    for (int rr = 0; rr < POLICY_N; rr += 32) {
        const int idx = ((start + rr) & (POLICY_N - 1)) + own;
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = wn[i];
        if (rr + 32 < POLICY_N) {
            const int nidx = ((start + rr + 32) & (POLICY_N - 1)) + own;
            const float4* wr = reinterpret_cast<const float4*>(H.wfc + (size_t)nidx * 128 + part * 16);
#pragma unroll
            for (int i = 0; i < 4; i++) wn[i] = __ldg(wr + i);
        }
        const float bias = __ldg(H.bfc + idx);
#pragma unroll
        for (int r = 0; r < kRM; r++) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a = fmaf(w[i].x, f[r][4 * i], a);
                a = fmaf(w[i].y, f[r][4 * i + 1], a);
                a = fmaf(w[i].z, f[r][4 * i + 2], a);
                a = fmaf(w[i].w, f[r][4 * i + 3], a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (part == 0) lall[r * POLICY_N + idx] = a + bias;
        }
    }
    __syncthreads();
    // statistics: warp w takes half h = w / 4 of leaf r = w % 4 (lane-strided, butterfly), the two halves combined in order
    const int r = wid & 3, hbase = (wid >> 2) * (POLICY_N / 2);
    const float* lr = lall + r * POLICY_N + hbase;
    float m = -3.0e38f;
    for (int i = lane; i < POLICY_N / 2; i += 32) m = lr[i] > m ? lr[i] : m;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, m, d);
        m = o > m ? o : m;
    }
    if (lane == 0) sm.pm[r][wid >> 2] = m;
    __syncthreads();
    const float M = sm.pm[r][0] > sm.pm[r][1] ? sm.pm[r][0] : sm.pm[r][1];
    float z = 0.f;
    for (int i = lane; i < POLICY_N / 2; i += 32) z = z + kvd_expf(lr[i] - M);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) z = z + __shfl_xor_sync(0xffffffffu, z, d);
    if (lane == 0) sm.pz[r][wid >> 2] = z;
    __syncthreads();
    if (threadIdx.x < kRM) {
        sm.mx[threadIdx.x] = sm.pm[threadIdx.x][0] > sm.pm[threadIdx.x][1] ? sm.pm[threadIdx.x][0] : sm.pm[threadIdx.x][1];
        sm.z[threadIdx.x] = sm.pz[threadIdx.x][0] + sm.pz[threadIdx.x][1];
    }
    for (int rr = 0; rr < nb; rr++) {
        const float g = cfg.dir_eps > 0.0f ? root_noise_sum_cta(cfg, A, sm.gs[rr], sm.red) : 0.f;
        if (threadIdx.x == 0) sm.gsum[rr] = g;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 2) mcts_eval_rootmix_kernel(MctsCfg cfg, MctsArrays A, const __nv_bfloat16* __restrict__ act,
                                                                HeadW H, uint32_t wave) {
    __shared__ RootMixSmem sm;
    __shared__ __align__(16) float swh[3 * 512];
    extern __shared__ float lall_dyn[];
    const int n_eval = (int)*A.n_eval;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s0 = blockIdx.x * kRM; s0 < n_eval; s0 += gridDim.x * kRM) {
        const int nb = n_eval - s0 < kRM ? n_eval - s0 : kRM;
        bool any_root = false;
        for (int r = 0; r < kRM; r++) {
            if (r < nb) {
                const int slot = s0 + r, gs = A.eval_game[slot];
                any_root |= A.pend_node[gs] == 0;
                kvn::head_features(act + (size_t)slot * 64 * H.C, H.C, H.wh, H.bh, sm.hp[r], sm.hv, swh);
                __syncthreads();
                const float v_white = kvn::value_mlp(sm.hv, H.w1, H.b1, H.w2, H.b2, sm.red);
                legal_logits(cfg, A, gs, H, sm.hp[r], sm.logits[r]);
                if (threadIdx.x == 0) {
                    sm.vw[r] = v_white;
                    sm.gs[r] = gs;
                }
                __syncthreads();
            } else {
                for (int i = threadIdx.x; i < 128; i += blockDim.x) sm.hp[r][i] = 0.f;   // defined input for the shared pass
            }
        }
        __syncthreads();
        if (any_root) rootmix_stats(cfg, A, H, sm, nb, lall_dyn);
        if (wid < nb) {   // warp r finishes leaf r
            const int gs = sm.gs[wid];
            const bool root = A.pend_node[gs] == 0;
            mcts_expand_warp(lane, cfg, A, gs, sm.logits[wid], sm.vw[wid], false, root ? sm.mx[wid] : 0.f, root ? sm.z[wid] : 0.f,
                             root ? sm.gsum[wid] : 0.f);
            if (cfg.cache_mask) cache_fill_warp(lane, cfg, A, wave, s0 + wid, sm.hp[wid], sm.vw[wid]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256, 2) mcts_late_rootmix_kernel(MctsCfg cfg, MctsArrays A, HeadW H) {
    __shared__ RootMixSmem sm;
    extern __shared__ float lall_dyn[];
    const int n_late = (int)*A.n_late;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int l0 = blockIdx.x * kRM; l0 < n_late; l0 += gridDim.x * kRM) {
        const int nb = n_late - l0 < kRM ? n_late - l0 : kRM;
        bool any_root = false;
        for (int r = 0; r < kRM; r++) {
            if (r < nb) {
                const int gs = A.late_game[l0 + r], src = A.late_src[l0 + r];
                any_root |= A.pend_node[gs] == 0;
                const float* f = src < 0 ? A.feat_game + (size_t)gs * FEAT : A.feat_slot + (size_t)src * FEAT;
                for (int i = threadIdx.x; i < FEAT; i += blockDim.x) sm.hp[r][i] = f[i];
                if (threadIdx.x == 0) sm.gs[r] = gs;
                __syncthreads();
                legal_logits(cfg, A, gs, H, sm.hp[r], sm.logits[r]);
                if (threadIdx.x == 0) sm.vw[r] = sm.hp[r][128];
            } else {
                for (int i = threadIdx.x; i < 128; i += blockDim.x) sm.hp[r][i] = 0.f;
            }
        }
        __syncthreads();
        if (any_root) rootmix_stats(cfg, A, H, sm, nb, lall_dyn);
        if (wid < nb) {
            const int gs = sm.gs[wid];
            const bool root = A.pend_node[gs] == 0;
            mcts_expand_warp(lane, cfg, A, gs, sm.logits[wid], sm.vw[wid], true, root ? sm.mx[wid] : 0.f, root ? sm.z[wid] : 0.f,
                             root ? sm.gsum[wid] : 0.f);
        }
        __syncthreads();
    }
}

__global__ void mcts_init_kernel(MctsCfg cfg, MctsArrays A, int G, const uint64_t* __restrict__ start, uint64_t id_base) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    GameHdr h;
    memset(&h, 0, sizeof(h));
    h.game_id = id_base + (uint64_t)g;
    A.hdr[g] = h;
    for (int j = 0; j < cfg.inflight; j++) A.pend_node[(size_t)g * cfg.inflight + j] = -1;
    for (int i = 0; i < LINE_WORDS; i++) {
        uint64_t v = start ? start[(size_t)g * LINE_WORDS + i] : 0ull;
        if (i >= 13) v = 0;
        A.root_line[(size_t)g * LINE_WORDS + i] = v;
    }
}

// status: [0] games done, [1] sum sims_done, [2] sum evals, [3] sum plies, [4] overflow count, [5] white wins,
// [6] black wins, [7] draws (among done), [8] expansions served by the cache, [9] games stopped by an illegal scripted move
__global__ void mcts_status_kernel(MctsArrays A, int G, unsigned long long* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const GameHdr h = A.hdr[g];
    if (h.done) atomicAdd(out + 0, 1ull);
    atomicAdd(out + 1, (unsigned long long)h.sims_done);
    atomicAdd(out + 2, (unsigned long long)h.n_evals);
    atomicAdd(out + 3, (unsigned long long)h.ply);
    if (h.overflow & HDR_OVERFLOW) atomicAdd(out + 4, 1ull);
    if (h.overflow & HDR_SCRIPT_MISS) atomicAdd(out + 9, 1ull);
    atomicAdd(out + 8, (unsigned long long)h.cache_hits);
    if (h.done && h.result > 0) atomicAdd(out + 5, 1ull);
    if (h.done && h.result < 0) atomicAdd(out + 6, 1ull);
    if (h.done && h.result == 0) atomicAdd(out + 7, 1ull);
}

// number of live games whose current move still has simulations to run (K > 1: waves are not one-per-simulation)
__global__ void mcts_unfinished_kernel(MctsCfg cfg, MctsArrays A, int G, unsigned int* out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const GameHdr h = A.hdr[g];
    if (!h.done && h.sims_done < cfg.sims) atomicAdd(out, 1u);
}

// Records of all games, in game order: lines [N][16] (bitboards, rest 0), move index, reward (self_play.py:245-250:
// white's perspective, win 1.0 / draw 0.2 / loss -1.0, same value on every ply), game id.
__global__ void __launch_bounds__(1024) mcts_record_offsets_kernel(MctsArrays A, int G, int rec_cap, int* offsets /*[G+1]*/) {
    // one CTA: every thread sums a contiguous chunk of games, a shared-memory scan orders the chunks (game order)
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (G + 1023) / 1024;
    const int lo = t * per, hi = lo + per < G ? lo + per : G;
    int acc = 0;
    for (int g = lo; g < hi; g++) {
        const int p = A.hdr[g].ply;
        acc += p < rec_cap ? p : rec_cap;
    }
    part[t] = acc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {   // inclusive Hillis-Steele scan
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int base = t ? part[t - 1] : 0;
    for (int g = lo; g < hi; g++) {
        offsets[g] = base;
        const int p = A.hdr[g].ply;
        base += p < rec_cap ? p : rec_cap;
    }
    if (t == 1023) offsets[G] = part[1023];
}
__global__ void mcts_record_gather_kernel(MctsArrays A, int G, int rec_cap, const int* __restrict__ offsets,
                                          uint64_t* __restrict__ out_lines, int32_t* __restrict__ out_move,
                                          float* __restrict__ out_reward, int32_t* __restrict__ out_game, int cap) {
    const int g = blockIdx.x;
    if (g >= G) return;
    const GameHdr h = A.hdr[g];
    const int np = h.ply < rec_cap ? h.ply : rec_cap;
    const float reward = h.result > 0 ? 1.0f : (h.result < 0 ? -1.0f : 0.2f);
    for (int i = threadIdx.x; i < np * 16; i += blockDim.x) {
        const int p = i >> 4, wd = i & 15;
        const int o = offsets[g] + p;
        if (o >= cap) continue;
        out_lines[(size_t)o * 16 + wd] = wd < 12 ? A.rec_line[((size_t)g * rec_cap + p) * 12 + wd] : 0ull;
        if (wd == 0) {
            out_move[o] = move_index(A.rec_move[(size_t)g * rec_cap + p]);
            out_reward[o] = reward;
            out_game[o] = g;
        }
    }
}

}  // namespace kv

using namespace kv;

struct kv_mcts {
    MctsCfg cfg;
    MctsArrays A;
    int G = 0;
    std::vector<void*> allocs;
    unsigned long long* d_status = nullptr;
    int* d_offsets = nullptr;
    uint32_t wave = 0;
    unsigned int* d_unfinished = nullptr;
    unsigned int* h_unfinished = nullptr;   // pinned
    long waves_run = 0;
    float* head_buf = nullptr;   // [G*K][192] head features of the wave's evaluated leaves (split evaluator)
    int eval_split = -1;         // -1 default (on unless KV_MCTS_EVAL_SPLIT=0), 0 one kernel per leaf, 1 split
    void* cache_mem = nullptr;
    size_t cache_slots = 0;
    // pipelined search: two game groups on two streams (see mcts_wave_group)
    int pipeline = -1;                 // -1 auto, 0 off, 1 on
    cudaStream_t side = nullptr;       // group 1's stream (group 0 runs on the caller's)
    cudaStream_t tower = nullptr;      // the tensor-core kernels of both groups, in issue order (highest priority)
    cudaEvent_t ev_stem[2] = {nullptr, nullptr};
    cudaEvent_t ev_tower[2] = {nullptr, nullptr}, ev_eval[2] = {nullptr, nullptr}, ev_late[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool tower_rec[2] = {false, false}, eval_rec[2] = {false, false}, late_rec[2] = {false, false};
    uint32_t last_wave[2] = {0, 0};
};

void kv_mcts_destroy(kv_ctx* ctx) {
    kv_mcts* m = ctx->mcts;
    if (!m) return;
    for (void* p : m->allocs) cudaFree(p);
    if (m->side) cudaStreamDestroy(m->side);
    if (m->tower) cudaStreamDestroy(m->tower);
    for (int i = 0; i < 2; i++) {
        if (m->ev_stem[i]) cudaEventDestroy(m->ev_stem[i]);
        if (m->ev_tower[i]) cudaEventDestroy(m->ev_tower[i]);
        if (m->ev_eval[i]) cudaEventDestroy(m->ev_eval[i]);
        if (m->ev_late[i]) cudaEventDestroy(m->ev_late[i]);
    }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->h_unfinished) cudaFreeHost(m->h_unfinished);
    if (m->cache_mem) cudaFree(m->cache_mem);
    delete m;
    ctx->mcts = nullptr;
}

template <class T>
static int dalloc(kv_ctx* ctx, kv_mcts* m, T** p, size_t count) {
    KV_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
    m->allocs.push_back(*p);
    return 0;
}

extern "C" {

int kv_mcts_create(kv_ctx* ctx, int n_games, int sims, int edges_per_node, int max_plies, int temp_plies, float c_puct,
                   float dir_alpha, float dir_eps, uint64_t seed, int eval_mode) {
    return kv_mcts_create_k(ctx, n_games, sims, edges_per_node, max_plies, temp_plies, c_puct, dir_alpha, dir_eps, seed,
                            eval_mode, 1);
}

int kv_mcts_create_k(kv_ctx* ctx, int n_games, int sims, int edges_per_node, int max_plies, int temp_plies, float c_puct,
                     float dir_alpha, float dir_eps, uint64_t seed, int eval_mode, int inflight) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_games < 1 || sims < 1 || max_plies < 1) return kv_fail_msg(ctx, "kv_mcts_create: bad sizes");
    if (inflight < 1 || inflight > 128) return kv_fail_msg(ctx, "kv_mcts_create: inflight must be in 1..128");
    if (sims >= (1 << 24)) return kv_fail_msg(ctx, "kv_mcts_create: sims must be < 2^24");
    if ((long long)n_games * inflight > (1ll << 24)) return kv_fail_msg(ctx, "kv_mcts_create: n_games * inflight too large");
    if (eval_mode == 1 && (!ctx->net || ctx->net->cap < n_games * inflight))
        return kv_fail_msg(ctx, "kv_mcts_create: network evaluator needs kv_net_create(max_boards >= n_games * inflight) first");
    kv_mcts_destroy(ctx);
    kv_mcts* m = new kv_mcts();
    ctx->mcts = m;
    m->G = n_games;
    MctsCfg& c = m->cfg;
    c.sims = sims;
    c.node_cap = sims;
    if (edges_per_node <= 0) edges_per_node = 64;   // 48 overflowed in 5 of 4 096 games at 800 sims (bench, round 2)
    c.edge_cap = sims * edges_per_node;
    if (c.edge_cap < MAX_MOVES) c.edge_cap = MAX_MOVES;
    c.temp_plies = temp_plies;
    c.max_plies = max_plies;
    c.rec_cap = max_plies;
    c.eval_mode = eval_mode;
    c.c_puct = c_puct;
    c.dir_alpha = dir_alpha;
    c.dir_eps = dir_eps;
    c.seed = seed;
    c.inflight = inflight;
    c.resign_thr = -0.7f;        // scripts/self_play.py:185
    c.resign_min_plies = 15;
    c.root_mix = sims == 1;      // no search: the move is sampled from the root priors, so they follow the reference's rule
    c.script_stride = 0;
    MctsArrays& A = m->A;
    const size_t G = (size_t)n_games;
    const size_t GS = G * (size_t)inflight;   // in-flight slots
    if (dalloc(ctx, m, &A.hdr, G)) return -1;
    if (dalloc(ctx, m, &A.root_line, G * 16)) return -1;
    if (dalloc(ctx, m, &A.node_line, G * c.node_cap * 16)) return -1;
    if (dalloc(ctx, m, &A.node_meta, G * c.node_cap)) return -1;
    if (dalloc(ctx, m, &A.eP, G * c.edge_cap)) return -1;
    if (dalloc(ctx, m, &A.eN, G * c.edge_cap)) return -1;
    if (dalloc(ctx, m, &A.eW, G * c.edge_cap)) return -1;
    if (dalloc(ctx, m, &A.eChild, G * c.edge_cap)) return -1;
    if (dalloc(ctx, m, &A.eMv, G * c.edge_cap)) return -1;
    if (dalloc(ctx, m, &A.path_edge, GS * (c.node_cap + 1))) return -1;
    if (dalloc(ctx, m, &A.path_node, GS * (c.node_cap + 1))) return -1;
    if (dalloc(ctx, m, &A.pend_node, GS)) return -1;
    if (dalloc(ctx, m, &A.pend_depth, GS)) return -1;
    if (dalloc(ctx, m, &A.pend_kind, GS)) return -1;
    if (dalloc(ctx, m, &A.n_eval, 4)) return -1;
    if (dalloc(ctx, m, &A.eval_game, GS)) return -1;
    if (dalloc(ctx, m, &A.eval_lines, (GS + 1) * 16)) return -1;
    if (dalloc(ctx, m, &A.rec_line, G * c.rec_cap * 12)) return -1;
    if (dalloc(ctx, m, &A.rec_move, G * c.rec_cap)) return -1;
    if (dalloc(ctx, m, &A.eval_centry, GS)) return -1;
    if (dalloc(ctx, m, &A.eval_hash, GS)) return -1;
    if (dalloc(ctx, m, &A.n_late, 4)) return -1;
    if (dalloc(ctx, m, &A.late_game, GS)) return -1;
    if (dalloc(ctx, m, &A.late_src, GS)) return -1;
    if (dalloc(ctx, m, &A.feat_game, GS * FEAT)) return -1;
    if (dalloc(ctx, m, &A.feat_slot, GS * FEAT)) return -1;
    if (eval_mode == 1 && dalloc(ctx, m, &m->head_buf, GS * HEADF)) return -1;
    if (dalloc(ctx, m, &m->d_unfinished, 4)) return -1;
    KV_CUDA(ctx, cudaMallocHost((void**)&m->h_unfinished, 16));
    A.cache = nullptr;
    c.cache_mask = 0;
    KV_CUDA(ctx, cudaMemset(A.n_late, 0, 16));
    if (dalloc(ctx, m, &m->d_status, 16)) return -1;
    if (dalloc(ctx, m, &m->d_offsets, G + 1)) return -1;
    KV_CUDA(ctx, cudaMemset(A.n_eval, 0, 16));
    return 0;
}

// start lines: d_start [n_games][16] or NULL for the initial position of GameState() (core/chessEngine.py:34-84)
int kv_mcts_reset(kv_ctx* ctx, const uint64_t* d_start, uint64_t game_id_base, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_reset: no search context");
    kv_mcts* m = ctx->mcts;
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t* tmp = nullptr;
    if (!d_start) {
        // the standard initial position as a board line, replicated
        static const uint64_t start[16] = {
            1ull << 60, 1ull << 59, (1ull << 56) | (1ull << 63), (1ull << 58) | (1ull << 61), (1ull << 57) | (1ull << 62),
            0xFFull << 48, 1ull << 4, 1ull << 3, (1ull << 0) | (1ull << 7), (1ull << 2) | (1ull << 5),
            (1ull << 1) | (1ull << 6), 0xFFull << 8,
            1ull | (64ull << 8) | (60ull << 16) | (4ull << 24), 0, 0, 0};
        std::vector<uint64_t> h((size_t)m->G * 16);
        for (int g = 0; g < m->G; g++) memcpy(&h[(size_t)g * 16], start, sizeof(start));
        KV_CUDA(ctx, cudaMalloc(&tmp, h.size() * 8));
        KV_CUDA(ctx, cudaMemcpyAsync(tmp, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
        KV_CUDA(ctx, cudaStreamSynchronize(st));
        d_start = tmp;
    }
    mcts_init_kernel<<<(m->G + 255) / 256, 256, 0, st>>>(m->cfg, m->A, m->G, d_start, game_id_base);
    KV_LAUNCH_CHECK(ctx);
    if (tmp) {
        KV_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
    return 0;
}

// One wave of the games [g0, g1) on stream st.  `grp` < 0: the whole context, one stream (the plain schedule).
// grp = 0 / 1: one group of the pipelined schedule —
//
//   stream A:  select A(w) stem A(w) ............ eval A(w) late A(w) select A(w+1) stem A(w+1) ............ eval A(w+1)
//   tower   :        ... B(w-1) | tower A(w)        | tower B(w)                         | tower A(w+1)       | ...
//   stream B:  eval B(w-1) late B(w-1) select B(w) stem B(w) ............ eval B(w) late B(w) select B(w+1) ...
//
// the tensor-core kernels of both groups run back to back on one high-priority stream, so the tensor pipe never waits for the tree kernels
// of the other group, which run on the CUDA cores of the same SMs meanwhile (the tower CTA leaves 18-33 KB of shared
// memory and most of the register file free).  Search results do not depend on the schedule: games are independent,
// and the evaluation cache — the only state the groups share — is transparent.  Its protocol between the groups:
// an entry under evaluation by the OTHER group's wave in flight (stamp == peer_wave) is followed like one of the own
// wave (late kernel after the other group's evaluator: ev_eval) and is never chosen as an eviction victim, so a
// fill never races with a claim of the same entry; the leader's features stay in feat_slot until the followers of
// both groups have read them (ev_late).
static int mcts_wave_group(kv_ctx* ctx, cudaStream_t st, int grp, int g0, int g1) {
    kv_mcts* m = ctx->mcts;
    const int K = m->cfg.inflight;
    const int G = g1 - g0, GS = G * K, s0 = g0 * K;
    const int grid = (G + kMW - 1) / kMW, grid_s = (GS + kMW - 1) / kMW;
    const uint32_t wave = ++m->wave;
    const int q = grp > 0 ? 1 : 0, peer = 1 - q;
    const bool piped = grp >= 0;
    // pipelined: grids sized to what fits beside one tower CTA per SM (16 K registers; 18 KB of shared memory next to the
    // halo-operand tower, 33 KB next to the 9-fetch one):
    // 2 selection CTAs (64 threads x 128 registers), 1 evaluator CTA (256 x 40), 3 late CTAs (128 x 40), 2 backup CTAs
    const int sms = ctx->sm_count;
    auto imin = [](int a, int b) { return a < b ? a : b; };
    // the group's view of the arrays: own counters and queues, global feat_slot
    MctsArrays A = m->A;
    A.n_eval += q;
    A.n_late += q;
    A.eval_game += s0;
    A.eval_lines += (size_t)s0 * LINE_WORDS;
    A.eval_centry += s0;
    A.eval_hash += s0;
    A.late_game += s0;
    A.late_src += s0;
    A.slot_base = s0;
    A.peer_wave = piped ? m->last_wave[peer] : 0;
    if (piped) m->last_wave[q] = wave;
    const bool cache = m->cfg.cache_mask != 0;
    KV_CUDA(ctx, cudaMemsetAsync(A.n_eval, 0, sizeof(uint32_t), st));
    if (cache) KV_CUDA(ctx, cudaMemsetAsync(A.n_late, 0, sizeof(uint32_t), st));
    {
        KvTimed t_(ctx, KVK_MCTS_SELECT, st);
        if (piped) mcts_select_kernel<kMWP><<<imin((G + kMWP - 1) / kMWP, 2 * sms), kMWP * 32, 0, st>>>(m->cfg, A, g0, g1, wave);
        else mcts_select_kernel<kMW><<<grid, kMW * 32, 0, st>>>(m->cfg, A, g0, g1, wave);
    }
    KV_LAUNCH_CHECK(ctx);
    auto wait_peer = [&](cudaEvent_t* ev, bool* rec) -> int {
        if (piped && rec[peer]) KV_CUDA(ctx, cudaStreamWaitEvent(st, ev[peer], 0));
        return 0;
    };
    auto mark = [&](cudaEvent_t* ev, bool* rec) -> int {
        if (piped) {
            KV_CUDA(ctx, cudaEventRecord(ev[q], st));
            rec[q] = true;
        }
        return 0;
    };
    if (m->cfg.eval_mode == 0) {
        // no tower here: ev_tower stands for "selection done", and the evaluator waits for the other group's, which
        // keeps the order the tower alternation gives the network path (a fill never meets a selection that does not
        // know its wave as peer_wave)
        if (int rc = wait_peer(m->ev_tower, m->tower_rec)) return rc;
        if (int rc = mark(m->ev_tower, m->tower_rec)) return rc;
        if (cache)
            if (int rc = wait_peer(m->ev_late, m->late_rec)) return rc;   // feat_slot of this group is rewritten now
        {
            KvTimed t_(ctx, KVK_MCTS_EXPAND, st);
            mcts_hash_eval_kernel<<<grid_s, kMW * 32, 0, st>>>(m->cfg, A, wave);
        }
        KV_LAUNCH_CHECK(ctx);
        if (cache) {
            if (int rc = mark(m->ev_eval, m->eval_rec)) return rc;
            if (int rc = wait_peer(m->ev_eval, m->eval_rec)) return rc;   // followers of the other group's leaders
            mcts_hash_late_kernel<<<grid_s, kMW * 32, 0, st>>>(m->cfg, A);
            KV_LAUNCH_CHECK(ctx);
            if (int rc = mark(m->ev_late, m->late_rec)) return rc;
        }
    } else {
        int fb = 0;
        kv_net* net = ctx->net;
        if (int rc = kv_net_tower(ctx, A.eval_lines, GS, st, &fb, -1, reinterpret_cast<const int*>(A.n_eval), s0,
                                  piped ? m->tower : nullptr, piped ? m->ev_stem[q] : nullptr, piped ? sms : 0))
            return rc;
        if (piped) {   // the group's stream continues when its tower is through
            KV_CUDA(ctx, cudaEventRecord(m->ev_tower[q], m->tower));
            m->tower_rec[q] = true;
            KV_CUDA(ctx, cudaStreamWaitEvent(st, m->ev_tower[q], 0));
        }
        const int cmax = net->C > net->C1 ? net->C : net->C1;
        const __nv_bfloat16* act = net->act[fb] + (size_t)s0 * 64 * cmax;
        HeadW H{net->wh, net->bh, net->wfc, net->bfc, net->w1, net->b1, net->w2, net->b2, net->C};
        if (cache)
            if (int rc = wait_peer(m->ev_late, m->late_rec)) return rc;
        {
            KvTimed t_(ctx, KVK_MCTS_EXPAND, st);
            static const int split_env = [] {
                const char* e = getenv("KV_MCTS_EVAL_SPLIT");
                return e ? atoi(e) : 0;   // opt-in: it shortens the evaluator by 10 % and the step by nothing (power cap)
            }();
            const bool split = (m->eval_split >= 0 ? m->eval_split : split_env) != 0 && !piped;
            if (m->cfg.root_mix) {
                mcts_eval_rootmix_kernel<<<piped ? imin((GS + kRM - 1) / kRM, sms) : (GS + kRM - 1) / kRM, 256, kRootMixDyn, st>>>(m->cfg, A, act, H, wave);
            } else if (split) {
                float* hb = m->head_buf + (size_t)s0 * HEADF;
                mcts_head_kernel<<<GS, 256, 0, st>>>(A, act, H, hb);
                mcts_eval_batch_kernel<<<(GS + kEB - 1) / kEB, 256, 0, st>>>(m->cfg, A, hb, H, wave);
            } else {
                mcts_eval_net_kernel<<<piped ? imin(GS, sms) : GS, 256, 0, st>>>(m->cfg, A, act, H, wave);
            }
        }
        KV_LAUNCH_CHECK(ctx);
        if (cache) {
            if (int rc = mark(m->ev_eval, m->eval_rec)) return rc;
            if (int rc = wait_peer(m->ev_eval, m->eval_rec)) return rc;
            {
                KvTimed t_(ctx, KVK_MCTS_EXPAND, st);
                if (m->cfg.root_mix)
                    mcts_late_rootmix_kernel<<<piped ? imin((GS + kRM - 1) / kRM, sms) : (GS + kRM - 1) / kRM, 256, kRootMixDyn, st>>>(m->cfg, A, H);
                else mcts_late_net_kernel<<<piped ? imin(GS, 3 * sms) : GS, 128, 0, st>>>(m->cfg, A, H);
            }
            KV_LAUNCH_CHECK(ctx);
            if (int rc = mark(m->ev_late, m->late_rec)) return rc;
        }
    }
    if (K > 1) {
        KvTimed t_(ctx, KVK_MCTS_EXPAND, st);
        mcts_backup_kernel<<<piped ? imin(grid, 2 * sms) : grid, kMW * 32, 0, st>>>(m->cfg, A, g0, g1);
        KV_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// Two groups only on request (kv_mcts_set_pipeline(1) or KV_MCTS_PIPELINE=1): measured on B200 the overlap is worth
// +1 % at most (4 096 games x 800 simulations: 3 985 ms against 4 029 ms per move) — the step is bound by the power cap,
// not by SM idle time, so hiding the tree kernels under the tower buys almost nothing.
// The split is a multiple of 8 games so that group 1's activations start on a tile boundary.
static int g_attrs_state = -1;   // carve-out preference last applied to the tree kernels (per process): 0 default, 1 max shared

static bool mcts_piped(const kv_ctx* ctx) {
    const kv_mcts* m = ctx->mcts;
    if (m->G < 16) return false;
    if (m->pipeline >= 0) return m->pipeline != 0;
    static const int env = [] {
        const char* e = getenv("KV_MCTS_PIPELINE");
        return e ? atoi(e) : 0;
    }();
    return env != 0 && m->cfg.eval_mode == 1 && ctx->net && (long long)m->G * m->cfg.inflight >= 2048;
}

static int mcts_pipe_setup(kv_ctx* ctx) {
    kv_mcts* m = ctx->mcts;
    if (m->side) return 0;
    int lo = 0, hi = 0;
    KV_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
    KV_CUDA(ctx, cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
    KV_CUDA(ctx, cudaStreamCreateWithPriority(&m->tower, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; i++) {
        KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_stem[i], cudaEventDisableTiming));
        KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_tower[i], cudaEventDisableTiming));
        KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_eval[i], cudaEventDisableTiming));
        KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_late[i], cudaEventDisableTiming));
    }
    KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    KV_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    return 0;
}

// n_waves waves for every game.  Pipelined: fork the side stream off the caller's, alternate the groups, join.
static int mcts_run_waves(kv_ctx* ctx, int n_waves, cudaStream_t st) {
    kv_mcts* m = ctx->mcts;
    if (n_waves <= 0) return 0;
    {
        // pipelined: the tree kernels share SMs with the tower CTAs (198-210 KB of dynamic shared memory), so they ask for the
        // same carve-out; otherwise the default (a large L1 keeps value_fc1's 131 KB resident for the evaluator CTAs)
        const int want = mcts_piped(ctx) ? 1 : 0;
        if (g_attrs_state != want) {
            const int carve = want ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
            cudaFuncSetAttribute(mcts_select_kernel<kMW>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(mcts_select_kernel<kMWP>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(mcts_eval_net_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(mcts_late_net_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(mcts_backup_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(mcts_eval_rootmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRootMixDyn);
            cudaFuncSetAttribute(mcts_late_rootmix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRootMixDyn);
            cudaGetLastError();
            g_attrs_state = want;
        }
    }
    m->waves_run += n_waves;
    if (!mcts_piped(ctx)) {
        for (int i = 0; i < n_waves; i++)
            if (int rc = mcts_wave_group(ctx, st, -1, 0, m->G)) return rc;
        return 0;
    }
    if (int rc = mcts_pipe_setup(ctx)) return rc;
    const int gh = ((m->G / 2) + 7) & ~7;
    KV_CUDA(ctx, cudaEventRecord(m->ev_fork, st));
    KV_CUDA(ctx, cudaStreamWaitEvent(m->side, m->ev_fork, 0));
    for (int i = 0; i < 2; i++) m->tower_rec[i] = m->eval_rec[i] = m->late_rec[i] = false;
    m->last_wave[0] = m->last_wave[1] = 0;
    for (int i = 0; i < n_waves; i++) {
        if (int rc = mcts_wave_group(ctx, st, 0, 0, gh)) return rc;
        if (int rc = mcts_wave_group(ctx, m->side, 1, gh, m->G)) return rc;
    }
    KV_CUDA(ctx, cudaEventRecord(m->ev_join, m->side));
    KV_CUDA(ctx, cudaStreamWaitEvent(st, m->ev_join, 0));
    return 0;
}

// live games that still owe simulations for the current move (one small kernel + a 4-byte read; synchronises)
static int mcts_unfinished(kv_ctx* ctx, cudaStream_t st, unsigned int* out) {
    kv_mcts* m = ctx->mcts;
    KV_CUDA(ctx, cudaMemsetAsync(m->d_unfinished, 0, sizeof(unsigned int), st));
    mcts_unfinished_kernel<<<(m->G + 255) / 256, 256, 0, st>>>(m->cfg, m->A, m->G, m->d_unfinished);
    KV_LAUNCH_CHECK(ctx);
    KV_CUDA(ctx, cudaMemcpyAsync(m->h_unfinished, m->d_unfinished, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    KV_CUDA(ctx, cudaStreamSynchronize(st));
    *out = *m->h_unfinished;
    return 0;
}

// Evaluation cache of 2^log2_slots entries x 640 B (0 = disable).  Keyed by the 12 bitboards (the network's whole
// input); search results are bit-identical with the cache on or off, only the number of tower evaluations changes.
int kv_mcts_enable_cache(kv_ctx* ctx, int log2_slots) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_enable_cache: no search context");
    kv_mcts* m = ctx->mcts;
    KV_CUDA(ctx, cudaDeviceSynchronize());
    if (m->cache_mem) cudaFree(m->cache_mem);
    m->cache_mem = nullptr;
    m->A.cache = nullptr;
    m->cfg.cache_mask = 0;
    m->cache_slots = 0;
    if (log2_slots <= 0) return 0;
    if (log2_slots < 8 || log2_slots > 28) return kv_fail_msg(ctx, "kv_mcts_enable_cache: log2_slots must be in 8..28");
    const size_t slots = (size_t)1 << log2_slots;
    KV_CUDA(ctx, cudaMalloc(&m->cache_mem, slots * sizeof(CacheEntry)));
    KV_CUDA(ctx, cudaMemset(m->cache_mem, 0, slots * sizeof(CacheEntry)));
    m->A.cache = reinterpret_cast<CacheEntry*>(m->cache_mem);
    m->cfg.cache_mask = (uint32_t)(slots - 1);
    m->cache_slots = slots;
    return 0;
}

// Must be called when the network weights change (cached features belong to the old weights).
int kv_mcts_cache_clear(kv_ctx* ctx, void* stream) {
    if (!ctx || !ctx->mcts) return 0;
    kv_mcts* m = ctx->mcts;
    if (m->cache_mem) KV_CUDA(ctx, cudaMemsetAsync(m->cache_mem, 0, m->cache_slots * sizeof(CacheEntry), (cudaStream_t)stream));
    return 0;
}

int kv_mcts_run_sims(kv_ctx* ctx, int n_waves, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_run_sims: no search context");
    return mcts_run_waves(ctx, n_waves, (cudaStream_t)stream);
}

// Pipelined search on two game groups: mode -1 default (off unless KV_MCTS_PIPELINE=1 is set), 0 off, 1 on (any
// evaluator, at least 16 games).  Search results are identical either way.
int kv_mcts_set_pipeline(kv_ctx* ctx, int mode) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_set_pipeline: no search context");
    if (mode < -1 || mode > 1) return kv_fail_msg(ctx, "kv_mcts_set_pipeline: mode must be -1, 0 or 1");
    ctx->mcts->pipeline = mode;
    return 0;
}

// Resignation rule of the game loop (scripts/self_play.py:184-189).  Defaults = the reference's (-0.7, 15);
// min_plies < 0 switches it off.
int kv_mcts_set_resign(kv_ctx* ctx, float threshold, int min_plies) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_set_resign: no search context");
    ctx->mcts->cfg.resign_thr = threshold;
    ctx->mcts->cfg.resign_min_plies = min_plies;
    return 0;
}

// Root prior rule: 0 = softmax + Dirichlet noise over the legal moves, 1 = the reference's mixing over all 4096 indices
// followed by the legal renormalisation (scripts/self_play.py:150-167), -1 = default (1 when sims == 1, else 0).
int kv_mcts_set_root_mix(kv_ctx* ctx, int mode) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_set_root_mix: no search context");
    if (mode < -1 || mode > 1) return kv_fail_msg(ctx, "kv_mcts_set_root_mix: mode must be -1, 0 or 1");
    ctx->mcts->cfg.root_mix = mode < 0 ? (ctx->mcts->cfg.sims == 1) : mode;
    return 0;
}

// Scripted play: d_moves [n_games][stride] move words (0xFFFF = choose as usual) and / or d_values [n_games][stride]
// (NaN = the evaluator's value) override, per game and ply, the move the game loop plays and the value its resign rule
// sees.  The arrays stay caller-owned and must outlive their use; (NULL, NULL, 0) clears the script.
int kv_mcts_set_script(kv_ctx* ctx, const uint16_t* d_moves, const float* d_values, int stride) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_set_script: no search context");
    if (stride < 0 || (stride == 0 && (d_moves || d_values))) return kv_fail_msg(ctx, "kv_mcts_set_script: bad stride");
    kv_mcts* m = ctx->mcts;
    m->A.script_move = stride ? d_moves : nullptr;
    m->A.script_val = stride ? d_values : nullptr;
    m->cfg.script_stride = (d_moves || d_values) ? stride : 0;
    return 0;
}

// Evaluator schedule of the tower path: 1 = head-features kernel + batched finish kernel, 0 = one kernel per leaf, -1 =
// default (0 unless KV_MCTS_EVAL_SPLIT=1).  Bit-identical results.
int kv_mcts_set_eval_split(kv_ctx* ctx, int mode) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_set_eval_split: no search context");
    if (mode < -1 || mode > 1) return kv_fail_msg(ctx, "kv_mcts_set_eval_split: mode must be -1, 0 or 1");
    ctx->mcts->eval_split = mode;
    return 0;
}

int kv_mcts_finish_move(kv_ctx* ctx, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_finish_move: no search context");
    kv_mcts* m = ctx->mcts;
    KvTimed t_(ctx, KVK_MCTS_MISC, (cudaStream_t)stream);
    mcts_finish_move_kernel<<<(m->G + kMW - 1) / kMW, kMW * 32, 0, (cudaStream_t)stream>>>(m->cfg, m->A, m->G);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

// one move for every live game: cfg.sims waves, then pick / record / play
int kv_mcts_run_move(kv_ctx* ctx, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_run_move: no search context");
    kv_mcts* m = ctx->mcts;
    const int K = m->cfg.inflight, S = m->cfg.sims;
    if (K == 1) {
        if (int rc = kv_mcts_run_sims(ctx, S, stream)) return rc;   // one simulation per game and wave
    } else {
        // wave 1 expands the root alone; afterwards at most K simulations per wave, fewer when selections collide
        // with leaves still under evaluation: run the minimum, then poll until every game has its S simulations
        int n = 1 + (S - 1 + K - 1) / K;
        unsigned int left = 1;
        for (int guard = 0; guard < 4 * S + 8 && left; guard++) {
            if (int rc = kv_mcts_run_sims(ctx, n, stream)) return rc;
            if (int rc = mcts_unfinished(ctx, (cudaStream_t)stream, &left)) return rc;
            n = 2;
        }
        if (left) return kv_fail_msg(ctx, "kv_mcts_run_move: games still owe simulations after the wave budget");
    }
    return kv_mcts_finish_move(ctx, stream);
}

int kv_mcts_status(kv_ctx* ctx, uint64_t* h_out8, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_status: no search context");
    kv_mcts* m = ctx->mcts;
    cudaStream_t st = (cudaStream_t)stream;
    KV_CUDA(ctx, cudaMemsetAsync(m->d_status, 0, 16 * sizeof(unsigned long long), st));
    mcts_status_kernel<<<(m->G + 255) / 256, 256, 0, st>>>(m->A, m->G, m->d_status);
    KV_LAUNCH_CHECK(ctx);
    KV_CUDA(ctx, cudaMemcpyAsync(h_out8, m->d_status, 10 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    KV_CUDA(ctx, cudaStreamSynchronize(st));
    return 0;
}

// test hook: root edges of one game (host buffers of 256 entries): move words, visits, W, P; info = {n, nodes, edges, ply}
int kv_mcts_read_root(kv_ctx* ctx, int game, uint16_t* h_moves, uint32_t* h_N, float* h_W, float* h_P, int32_t* h_info4) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_read_root: no search context");
    kv_mcts* m = ctx->mcts;
    KV_CUDA(ctx, cudaDeviceSynchronize());
    GameHdr h;
    KV_CUDA(ctx, cudaMemcpy(&h, m->A.hdr + game, sizeof(h), cudaMemcpyDeviceToHost));
    NodeMeta nm;
    KV_CUDA(ctx, cudaMemcpy(&nm, m->A.node_meta + (size_t)game * m->cfg.node_cap, sizeof(nm), cudaMemcpyDeviceToHost));
    int n = (h.n_nodes && !(nm.ne_term & NODE_TERM)) ? (nm.ne_term & 0xFFFF) : 0;
    h_info4[0] = n;
    h_info4[1] = h.n_nodes;
    h_info4[2] = h.n_edges;
    h_info4[3] = h.ply;
    if (n) {
        const size_t e0 = (size_t)game * m->cfg.edge_cap + nm.first_edge;
        KV_CUDA(ctx, cudaMemcpy(h_moves, m->A.eMv + e0, n * 2, cudaMemcpyDeviceToHost));
        KV_CUDA(ctx, cudaMemcpy(h_N, m->A.eN + e0, n * 4, cudaMemcpyDeviceToHost));
        KV_CUDA(ctx, cudaMemcpy(h_W, m->A.eW + e0, n * 4, cudaMemcpyDeviceToHost));
        KV_CUDA(ctx, cudaMemcpy(h_P, m->A.eP + e0, n * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// test hook: what the oracle needs to replay a search "given identical net outputs": per node the evaluator's value
// and first edge, per edge the prior.  Host buffers sized node_cap / node_cap / edge_cap.
int kv_mcts_dump_tree(kv_ctx* ctx, int game, float* h_node_val, int32_t* h_node_first, float* h_edge_P,
                      uint64_t* h_root_line16) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_dump_tree: no search context");
    kv_mcts* m = ctx->mcts;
    KV_CUDA(ctx, cudaDeviceSynchronize());
    std::vector<NodeMeta> nm(m->cfg.node_cap);
    KV_CUDA(ctx, cudaMemcpy(nm.data(), m->A.node_meta + (size_t)game * m->cfg.node_cap, nm.size() * sizeof(NodeMeta),
                            cudaMemcpyDeviceToHost));
    for (int i = 0; i < m->cfg.node_cap; i++) {
        h_node_val[i] = nm[i].val;
        h_node_first[i] = nm[i].first_edge;
    }
    KV_CUDA(ctx, cudaMemcpy(h_edge_P, m->A.eP + (size_t)game * m->cfg.edge_cap, (size_t)m->cfg.edge_cap * 4,
                            cudaMemcpyDeviceToHost));
    if (h_root_line16)
        KV_CUDA(ctx, cudaMemcpy(h_root_line16, m->A.root_line + (size_t)game * 16, 128, cudaMemcpyDeviceToHost));
    return 0;
}

// current position of every game: d_lines [n_games][16]
int kv_mcts_get_roots(kv_ctx* ctx, uint64_t* d_lines, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_get_roots: no search context");
    KV_CUDA(ctx, cudaMemcpyAsync(d_lines, ctx->mcts->A.root_line, (size_t)ctx->mcts->G * 128, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return 0;
}

int kv_mcts_geometry(kv_ctx* ctx, int32_t* out4) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_geometry: no search context");
    out4[0] = ctx->mcts->G;
    out4[1] = ctx->mcts->cfg.node_cap;
    out4[2] = ctx->mcts->cfg.edge_cap;
    out4[3] = ctx->mcts->cfg.rec_cap;
    return 0;
}

/* waves launched since kv_mcts_create (K > 1: a move takes a data-dependent number of waves) */
int64_t kv_mcts_waves(kv_ctx* ctx) {
    return (ctx && ctx->mcts) ? (int64_t)ctx->mcts->waves_run : -1;
}

// Game records in game order (scripts/self_play.py:253 tuple fields, packed): returns the record count in *h_count.
// d_lines [cap][16] (feed to kv_encode for the reference's float planes), d_move [cap] policy index, d_reward [cap],
// d_game [cap] local game index.
int kv_mcts_records(kv_ctx* ctx, uint64_t* d_lines, int32_t* d_move, float* d_reward, int32_t* d_game, int cap,
                    int32_t* h_count, void* stream) {
    if (!ctx || !ctx->mcts) return kv_fail_msg(ctx, "kv_mcts_records: no search context");
    kv_mcts* m = ctx->mcts;
    cudaStream_t st = (cudaStream_t)stream;
    mcts_record_offsets_kernel<<<1, 1024, 0, st>>>(m->A, m->G, m->cfg.rec_cap, m->d_offsets);
    KV_LAUNCH_CHECK(ctx);
    int total = 0;
    KV_CUDA(ctx, cudaMemcpyAsync(&total, m->d_offsets + m->G, sizeof(int), cudaMemcpyDeviceToHost, st));
    KV_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_count) *h_count = total;
    if (d_lines && cap > 0) {
        mcts_record_gather_kernel<<<m->G, 128, 0, st>>>(m->A, m->G, m->cfg.rec_cap, m->d_offsets, d_lines, d_move, d_reward,
                                                        d_game, cap);
        KV_LAUNCH_CHECK(ctx);
    }
    return 0;
}

}  // extern "C"
