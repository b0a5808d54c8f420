// kv_net.cu — the policy/value network (reference: ai/model.py:8-77) for sm_100a.
//
//   stem      conv1 12->C1 + bn1 + relu, fused with board encoding (ai/ai.py:17-41): the input planes are
//             one-hot, so a 3x3 conv is at most 9 table rows added per output pixel; reads 96 B of bitboards
//             per board instead of 3 072 B of fp32 planes.  CUDA cores (0.11 % of the FLOPs).
//   tower     every other 3x3 conv (conv2 and the residual blocks, 99.8 % of the FLOPs) is ONE kernel:
//             a persistent, warp-specialised implicit GEMM on the 5th-gen tensor cores.
//               A  activations, NHWC bf16.  One TMA 4-D box {64 ch, 8, 8, 2 boards} per (filter tap, 64-channel
//                  block) with start coordinates (c0, dx, dy, b0), dx,dy in {-1,0,1}: the TMA unit zero-fills
//                  out-of-board pixels, i.e. the conv padding, so there is no im2col anywhere.  The box lands
//                  in shared memory as 128 rows x 128 B in the canonical K-major SWIZZLE_128B UMMA layout.
//               B  BN-folded weights [Cout][9*Cin] bf16 (K-major), TMA 2-D boxes {64, 256}.
//               D  128 x 256 fp32 accumulator in TMEM, double-buffered (2 x 256 of the 512 columns) so the
//                  epilogue of tile i overlaps the main loop of tile i+1.
//             warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue
//             (tcgen05.ld -> +bias (+residual) -> relu -> bf16 -> global).
//   heads     policy 1x1 conv + bn + relu + FC(128->4096) and value 1x1 conv + bn + relu + FC + relu + FC + tanh
//             fused in one CUDA-core kernel per board (0.04 % of the FLOPs); in search mode only the legal
//             moves' logits are computed and soft-maxed (kv_mcts.cu).
// BatchNorm is folded with running statistics (eval mode, eps 1e-5), as the reference runs self-play.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kv_internal.h"
#include "kv_net.h"
#include "kv_heads.cuh"
#include "kv_umma.cuh"
#include "kv_tower_order.h"

using bf16 = __nv_bfloat16;

namespace kvn {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CONV_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int CONV_THREADS = 256;

struct ConvParams {
    const float* bias;
    const bf16* residual;
    bf16* out;
    int m_tiles, n_tiles, kb_per_tap, cout, m_valid, relu;
    const int* n_ptr;   // optional device-side board count (search waves): overrides m_tiles / m_valid
    int board_base;     // first board of this launch inside the activation buffers (out / residual are already offset)
};

// epilogue of one 32-column chunk of one output row: +bias (+residual) -> ReLU -> bf16 -> four 16 B stores
__device__ __forceinline__ void conv_epilogue_store(const ConvParams& P, const uint32_t (&v)[32], size_t rbase, int n_tile,
                                                    int c) {
    const float4* bp = reinterpret_cast<const float4*>(P.bias + n_tile * BN + c * 32);
    uint4 res[4];
    if (P.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(P.residual + rbase + c * 32);
#pragma unroll
        for (int j = 0; j < 4; j++) res[j] = __ldg(rp + j);
    }
    uint4 o[4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 b = __ldg(bp + j);
        float x0 = __uint_as_float(v[4 * j + 0]) + b.x, x1 = __uint_as_float(v[4 * j + 1]) + b.y;
        float x2 = __uint_as_float(v[4 * j + 2]) + b.z, x3 = __uint_as_float(v[4 * j + 3]) + b.w;
        if (P.residual) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[j >> 1]) + (j & 1) * 2;
            const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rw[0]);
            const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rw[1]);
            x0 += __low2float(r01); x1 += __high2float(r01);
            x2 += __low2float(r23); x3 += __high2float(r23);
        }
        if (P.relu) {
            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
        }
        const __nv_bfloat162 p01 = __floats2bfloat162_rn(x0, x1), p23 = __floats2bfloat162_rn(x2, x3);
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o[j >> 1]) + (j & 1) * 2;
        ow[0] = *reinterpret_cast<const uint32_t*>(&p01);
        ow[1] = *reinterpret_cast<const uint32_t*>(&p23);
    }
    uint4* op = reinterpret_cast<uint4*>(P.out + rbase + c * 32);
#pragma unroll
    for (int j = 0; j < 4; j++) op[j] = o[j];
}

// same, with the residual chunk already in registers (prefetched before the accumulator was ready)
__device__ __forceinline__ void conv_epilogue_store_pf(const ConvParams& P, const uint32_t (&v)[32], const uint4 (&res)[4],
                                                       size_t rbase, int col0) {
    const float4* bp = reinterpret_cast<const float4*>(P.bias + col0);
    uint4 o[4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 b = __ldg(bp + j);
        float x0 = __uint_as_float(v[4 * j + 0]) + b.x, x1 = __uint_as_float(v[4 * j + 1]) + b.y;
        float x2 = __uint_as_float(v[4 * j + 2]) + b.z, x3 = __uint_as_float(v[4 * j + 3]) + b.w;
        if (P.residual) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[j >> 1]) + (j & 1) * 2;
            const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rw[0]);
            const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rw[1]);
            x0 += __low2float(r01); x1 += __high2float(r01);
            x2 += __low2float(r23); x3 += __high2float(r23);
        }
        if (P.relu) {
            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
        }
        const __nv_bfloat162 p01 = __floats2bfloat162_rn(x0, x1), p23 = __floats2bfloat162_rn(x2, x3);
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o[j >> 1]) + (j & 1) * 2;
        ow[0] = *reinterpret_cast<const uint32_t*>(&p01);
        ow[1] = *reinterpret_cast<const uint32_t*>(&p23);
    }
    uint4* op = reinterpret_cast<uint4*>(P.out + rbase);
#pragma unroll
    for (int j = 0; j < 4; j++) op[j] = o[j];
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        kvu::prefetch_tmap(&tmA);
        kvu::prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            kvu::mbar_init(&full[s], 1);
            kvu::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            kvu::mbar_init(&tfull[a], 1);
            kvu::mbar_init(&tempty[a], 128);
        }
        kvu::fence_barrier_init();
    }
    if (warp == 2) kvu::tmem_alloc(tmem_slot, 512);
    kvu::tc_fence_before();
    __syncthreads();
    kvu::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (P.n_ptr) {
        const int nb = *P.n_ptr;
        P.m_tiles = (nb + 1) >> 1;
        P.m_valid = nb * 64;
    }
    const int total = P.m_tiles * P.n_tiles;
    const int ksteps = 9 * P.kb_per_tap;

    if (warp == 0) {
        // ---- TMA producer ---------------------------------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int m_tile = tile / P.n_tiles, n_tile = tile % P.n_tiles;
            for (int ks = 0; ks < ksteps; ks++) {
                const int tap = ks / P.kb_per_tap, kb = ks - tap * P.kb_per_tap;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                kvu::mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    kvu::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    kvu::tma_load_4d(sa, &tmA, &full[stage], kb * BK, dx, dy, P.board_base + m_tile * 2);
                    kvu::tma_load_2d(sa + A_BYTES, &tmB, &full[stage], ks * BK, n_tile * BN);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer -----------------------------------------------------------------------------
        constexpr uint32_t idesc = kvu::make_idesc_bf16(BM, BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            kvu::mbar_wait(&tempty[acc], acc_phase ^ 1);
            kvu::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
            for (int ks = 0; ks < ksteps; ks++) {
                kvu::mbar_wait(&full[stage], phase);
                kvu::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = kvu::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = kvu::make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = kvu::make_sw128_kmajor_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)   // +32 B per UMMA_K inside the 128 B swizzle atom
                        kvu::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0);
                    kvu::umma_commit(&empty[stage]);
                    if (ks == ksteps - 1) kvu::umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ---- epilogue: TMEM -> registers -> bias/residual/relu -> bf16 -> global ------------------------
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int m_tile = tile / P.n_tiles, n_tile = tile % P.n_tiles;
            kvu::mbar_wait(&tfull[acc], acc_phase);
            kvu::tc_fence_after();
            const int row = m_tile * BM + q * 32 + lane;
            const bool valid = row < P.m_valid;
            const size_t rbase = (size_t)row * P.cout + (size_t)n_tile * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                kvu::tmem_ld_32x32(tmem_base + (uint32_t)acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                kvu::tmem_ld_wait();
                if (valid) conv_epilogue_store(P, v, rbase, n_tile, c);
            }
            kvu::tc_fence_before();
            kvu::mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    kvu::tc_fence_before();
    __syncthreads();
    if (warp == 2) kvu::tmem_dealloc(tmem_base, 512);
}

// ---- the same implicit GEMM on CTA pairs (cta_group::2): M = 256 rows (4 boards) x N = 256 per cluster-tile ----------
// Each CTA TMA-loads its own 128 activation rows and HALF of the weight tile (128 of the 256 output channels); the
// tensor core of the pair reads both halves, so every SM fetches 32 KB instead of 48 KB per k-step and the same
// 192 KB of shared memory holds 6 stages instead of 4 (50 % more latency cover for the L2 -> SMEM stream).
// Leader CTA (rank 0): its `full` barriers collect the bytes of both CTAs, its warp 1 issues the MMAs and commits
// (multicast) to both CTAs' `empty` / `tfull` barriers; both epilogues arrive on the leader's `tempty`.
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = (BN / 2) * BK * 2, STAGE2_BYTES = A_BYTES + B2_BYTES;
constexpr int CONV2_SMEM = STAGES2 * STAGE2_BYTES + 1024 + 256;

constexpr int CONV2_THREADS = 384;   // warps 0-2 producer / MMA / TMEM, warp 3 idle, warps 4-11 epilogue

// ---- the whole tower as ONE launch: a dependency-driven schedule over (layer, board tile, channel tile) tasks ------
// A 3x3 convolution is board-local, so tile (l, m, n) — layer l, boards [4m, 4m+4), output channels [256n, 256n+256) —
// needs exactly the two tiles (l-1, m, 0..1).  Tasks are numbered layer-major, t = (l * M + m) * NT + n, and CTA pair
// c runs t = c, c + pairs, ... in order: with >= 74 tiles per layer a task's inputs were finished about a layer's worth
// of time ago, so the waits below almost never spin, there is NO grid-wide barrier and no per-layer tail (the eleven
// tile-granular tails of the per-layer launches — half a round each on average, 3 % of a 2 500-board tower — shrink
// to one).  done[l][m] counts the epilogue warps that have stored their part of (l, m, .): channel tiles x 2 CTAs x
// 8 warps (32 for the 512-channel tower).  Writers: st.global -> fence.proxy.async (the readers are TMA loads, the async proxy) -> __threadfence
// -> red.release; readers: ld.acquire -> fence.proxy.async -> TMA.  Deadlock-free: a task waits only for tasks with a
// smaller number, every pair runs its tasks in increasing order, and all pairs are resident (grid <= 148 CTAs, 1 / SM).
// The spin is bounded and traps (a protocol bug must not hang the GPU box).
struct TowerLayerDev {
    CUtensorMap wmap;        // this layer's weights, {64, 128} boxes (one CTA's half of the N tile)
    CUtensorMap wmap_q;      // {64, 64} boxes: a quarter of the N tile (4-CTA clusters: two CTAs multicast a half each)
    const float* bias;
    int in_buf, in_view;     // activation buffer 0..2 and view (0: C1 channels, 1: C channels) of the input
    int out_buf, res_buf;    // res_buf < 0: no residual
    int kb_per_tap, relu;
};
static_assert(sizeof(TowerLayerDev) % 64 == 0, "TowerLayerDev: the tensor map of every array element must stay 64 B aligned");

struct TowerActMaps {
    CUtensorMap m[3][2];   // boxes {64 ch, 8 x, 8 y, 2 boards}
    CUtensorMap h[3][2];   // boxes {64 ch, 8 x, 2 boards, 10 y} (halo variant)
};

struct TowerArgs {
    bf16* act[3];            // activation buffers, already offset to this launch's first board
    const TowerLayerDev* layers;
    uint32_t* done;          // [n_layers][m_stride], zeroed before the launch
    const int* n_ptr;        // optional device-side board count
    int n_layers, m_stride, n_boards, cout;
    int board_base[2];       // first board of this launch in the C1 / C view of the buffers
    int chunk_tiles;         // depth-first order: board tiles per chunk (0 = layer-major over all boards)
    TowerLayerDev single;    // layers == nullptr: the one layer of this launch (a convolution on caller-owned tensors)
    // hybrid launch: this kernel computes only a slice of the board tiles — part 1: tiles [0, Ma), part 2: [Ma, M), with
    // Ma = M * split_num / split_den rounded down to even (M may be a device-side count); part 0: everything
    int part, split_num, split_den;
};
__device__ __forceinline__ void tower_slice(const TowerArgs& T, int M, int& m_off, int& m_cnt) {
    tower_slice_tiles(T.part, T.split_num, T.split_den, M, m_off, m_cnt);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one lane polls, the warp follows; `need` = epilogue warps per board tile = channel tiles x 2 CTAs x 8 warps
__device__ __forceinline__ void tower_wait_tile(const uint32_t* flag, uint32_t need, int lane) {
    if (lane == 0) {
        uint32_t it = 0;
        while (ld_acquire_u32(flag) < need) {
            __nanosleep(64);
            if (++it > (1u << 24)) __trap();
        }
    }
    __syncwarp();
}

// HALO = true: the activation tile is fetched THREE times per 64-channel block instead of nine.  One TMA box per column
// shift dx, `{64 ch, 8 x, 2 boards, 10 y}` from (c0, dx, b0, -1), lands as rows ordered [y = -1..8][board][x] (20 groups of
// 8 rows x 128 B, the groups y = -1 and y = 8 zero-filled by the TMA unit like the out-of-board columns).  A row shift dy is
// then a 2 KB offset of the operand's start address — two whole 8-row swizzle atoms, so the SWIZZLE_128B pattern is
// preserved — and the three taps (dy = -1, 0, 1; dx) read the same 20 KB through descriptors at +0 / +2 KB / +4 KB.
// A stage = that box + the three weight tiles of the taps: 68 KB, three stages.  Per 64-channel block and CTA the
// L2 -> shared-memory stream is 60 + 144 KB instead of 144 + 144 KB.  TMEM lane r of a CTA is (y, board, x) =
// (r / 16, (r / 8) % 2, r % 8); the taps are accumulated in the order (dx, dy) instead of (dy, dx), so results agree
// with the other kernels to fp32 rounding, not bit for bit.
constexpr int STAGES_H = 3;
constexpr int AH_BYTES = 20 * 1024, STAGE_H_BYTES = AH_BYTES + 3 * B2_BYTES;
constexpr int TOWER_H_SMEM = STAGES_H * STAGE_H_BYTES + 1024 + 256;

// 128 registers (a few dozen bytes of epilogue spill): 384 threads x 128 leave a quarter of the register file — and 18-33 KB
// of shared memory — to the tree-search kernels that share the SM with this CTA in the pipelined search (kv_mcts.cu).
// With T.layers == nullptr the launch is ONE convolution described by T.single (the per-layer mode of the tower and the
// training-side convolutions on caller-owned tensors).
template <bool HALO>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(128)
tower_umma2_kernel(const __grid_constant__ TowerActMaps maps, const __grid_constant__ TowerArgs T) {
    constexpr int NSTAGE = HALO ? STAGES_H : STAGES2;
    constexpr int SBYTES = HALO ? STAGE_H_BYTES : STAGE2_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * SBYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kvu::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            kvu::mbar_init(&full[s], 1);
            kvu::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; a++) {
            kvu::mbar_init(&tfull[a], 1);
            kvu::mbar_init(&tempty[a], 512);   // 256 epilogue threads of each CTA
        }
        kvu::fence_barrier_init();
    }
    if (warp == 2) kvu::tmem_alloc2(tmem_slot, 512);
    kvu::tc_fence_before();
    __syncthreads();
    kvu::cluster_sync();
    kvu::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nb = T.n_ptr ? *T.n_ptr : T.n_boards;
    const int NT = T.cout / BN;
    int m_off, M;
    tower_slice(T, (nb + 3) >> 2, m_off, M);
    const int m_valid = nb * 64;
    TowerOrder ord;
    ord.init(M, NT, T.n_layers, T.chunk_tiles);
    const int total = ord.total;
    const uint32_t need = 16u * (uint32_t)NT;

    if (warp == 0) {
        // ---- TMA producer (both CTAs) ---------------------------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, m_tile, n_tile;
            ord.decode(t, l, m_tile, n_tile);
            m_tile += m_off;
            const TowerLayerDev* L = T.layers ? T.layers + l : &T.single;
            const int kb_per_tap = L->kb_per_tap, ksteps = 9 * kb_per_tap;
            [[maybe_unused]] const CUtensorMap* tmA = &maps.m[L->in_buf][L->in_view];
            const int b0 = T.board_base[L->in_view] + m_tile * 4 + (int)rank * 2;
            if (l > 0) {   // the input tile must be complete (all channels of these four boards)
                tower_wait_tile(T.done + (size_t)(l - 1) * T.m_stride + m_tile, need, lane);
                fence_proxy_async();
            }
            if constexpr (HALO) {
                const CUtensorMap* tmH = &maps.h[L->in_buf][L->in_view];
                for (int st3 = 0; st3 < 3 * kb_per_tap; st3++) {   // (channel block, dx)
                    const int kb = st3 / 3, dxi = st3 - kb * 3;
                    kvu::mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) {
                        uint8_t* sa = smem + stage * SBYTES;
                        if (rank == 0) kvu::mbar_arrive_expect_tx(&full[stage], 2 * SBYTES);
                        kvu::tma2_load_4d(sa, tmH, &full[stage], kb * BK, dxi - 1, b0, -1);
#pragma unroll
                        for (int dyi = 0; dyi < 3; dyi++)
                            kvu::tma2_load_2d(sa + AH_BYTES + dyi * B2_BYTES, &L->wmap, &full[stage],
                                              ((dyi * 3 + dxi) * kb_per_tap + kb) * BK, n_tile * BN + (int)rank * (BN / 2));
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            } else {
            for (int ks = 0; ks < ksteps; ks++) {
                const int tap = ks / kb_per_tap, kb = ks - tap * kb_per_tap;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                kvu::mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * SBYTES;
                    if (rank == 0) kvu::mbar_arrive_expect_tx(&full[stage], 2 * SBYTES);
                    kvu::tma2_load_4d(sa, tmA, &full[stage], kb * BK, dx, dy, b0);
                    kvu::tma2_load_2d(sa + A_BYTES, &L->wmap, &full[stage], ks * BK, n_tile * BN + (int)rank * (BN / 2));
                }
                __syncwarp();
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ---- MMA issuer (leader CTA only) -------------------------------------------------------------------
        constexpr uint32_t idesc = kvu::make_idesc_bf16(2 * BM, BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, m_tile, n_tile;
            ord.decode(t, l, m_tile, n_tile);
            m_tile += m_off;
            const int ksteps = 9 * (T.layers ? T.layers[l].kb_per_tap : T.single.kb_per_tap);
            kvu::mbar_wait(&tempty[acc], acc_phase ^ 1);
            kvu::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
            if constexpr (HALO) {
                const int nst = ksteps / 3;   // stages of this tile: (channel block, dx)
                for (int s3 = 0; s3 < nst; s3++) {
                    kvu::mbar_wait(&full[stage], phase);
                    kvu::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = kvu::smem_u32(smem + stage * SBYTES);
#pragma unroll
                        for (int dyi = 0; dyi < 3; dyi++) {
                            const uint64_t adesc = kvu::make_sw128_kmajor_desc(sa + dyi * 2048);
                            const uint64_t bdesc = kvu::make_sw128_kmajor_desc(sa + AH_BYTES + dyi * B2_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; k++)
                                kvu::umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (s3 | dyi | k) != 0);
                        }
                        kvu::umma2_commit_mc(&empty[stage]);
                        if (s3 == nst - 1) kvu::umma2_commit_mc(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            } else {
            for (int ks = 0; ks < ksteps; ks++) {
                kvu::mbar_wait(&full[stage], phase);
                kvu::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = kvu::smem_u32(smem + stage * SBYTES);
                    const uint64_t adesc = kvu::make_sw128_kmajor_desc(sa);
                    const uint64_t bdesc = kvu::make_sw128_kmajor_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        kvu::umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0);
                    kvu::umma2_commit_mc(&empty[stage]);
                    if (ks == ksteps - 1) kvu::umma2_commit_mc(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ---- epilogue (both CTAs, own 128 rows): 8 warps = 4 TMEM lane quarters x 2 column halves -------------------
        const int q = warp & 3, half = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, m_tile, n_tile;
            ord.decode(t, l, m_tile, n_tile);
            m_tile += m_off;
            const TowerLayerDev* L = T.layers ? T.layers + l : &T.single;
            ConvParams P;
            P.bias = L->bias;
            P.out = T.act[L->out_buf];
            P.residual = L->res_buf >= 0 ? T.act[L->res_buf] : nullptr;
            P.relu = L->relu;
            int row = m_tile * 2 * BM + (int)rank * BM + q * 32 + lane;
            if constexpr (HALO) {   // TMEM lane r = (y, board, x): r / 16, (r / 8) % 2, r % 8
                const int r = q * 32 + lane;
                row = m_tile * 2 * BM + (int)rank * BM + ((r >> 3) & 1) * 64 + (r >> 4) * 8 + (r & 7);
            }
            const bool valid = row < m_valid;
            const int colb = n_tile * BN + half * (BN / 2);
            const size_t rbase = (size_t)row * T.cout + (size_t)colb;
            uint4 res[4][4];
            if (P.residual) {
                // the residual is the output of layer l-2 for these boards: complete once (l-1, m, .) is, which the
                // accumulator below cannot precede anyway
                if (l > 0) tower_wait_tile(T.done + (size_t)(l - 1) * T.m_stride + m_tile, need, lane);
                if (valid) {
                    const uint4* rp = reinterpret_cast<const uint4*>(P.residual + rbase);
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int j = 0; j < 4; j++) res[c][j] = __ldcg(rp + c * 4 + j);
                }
            }
            kvu::mbar_wait(&tfull[acc], acc_phase);
            kvu::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t v[32];
                kvu::tmem_ld_32x32(tmem_base + (uint32_t)acc * BN + half * (BN / 2) + c * 32 + ((uint32_t)(q * 32) << 16), v);
                kvu::tmem_ld_wait();
                if (valid) conv_epilogue_store_pf(P, v, res[c], rbase + c * 32, colb + c * 32);
            }
            kvu::tc_fence_before();
            kvu::mbar_arrive_leader(&tempty[acc]);
            // publish this warp's part of the tile to the TMA loads of the next layer
            fence_proxy_async();
            __threadfence();
            __syncwarp();
            if (lane == 0 && T.done) red_release_add_u32(T.done + (size_t)l * T.m_stride + m_tile, 1u);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    kvu::tc_fence_before();
    __syncthreads();
    kvu::cluster_sync();   // the peer may still be signalling this CTA's barriers / reading its B half
    if (warp == 2) kvu::tmem_dealloc2(tmem_base, 512);
}

// ---- 4-CTA clusters: two CTA pairs on neighbouring board tiles share every weight tile ---------------------------------
// After the halo operand 70 % of the L2 -> shared-memory stream is weights, fetched once per board tile.  Here a cluster
// is two pairs (CTAs 0,1 and 2,3) working on board tiles 2s and 2s+1 of the same layer and channel tile; the half of the
// weight tile that CTA r needs is the one CTA r^2 needs too, so each of the two loads a QUARTER ({64 k, 64 rows} box)
// and multicasts it to both: every CTA still receives 48 KB of weights per stage but only 24 KB per CTA cross L2 -> SM
// (132 instead of 204 KB per channel block and CTA).  A stage slot is now written by loads of the other pair, so the
// `empty` barriers collect the commits of BOTH pairs' MMA issuers (count 2, commit multicast to all four CTAs); the
// `full` / `tfull` / `tempty` barriers stay per pair.  Tasks are (layer, pair of board tiles, channel tile).
template <int UNUSED = 0>
__global__ void __cluster_dims__(4, 1, 1) __maxnreg__(128)
tower_umma4_kernel(const __grid_constant__ TowerActMaps maps, const __grid_constant__ TowerArgs T) {
    constexpr int NSTAGE = STAGES_H;
    constexpr int SBYTES = STAGE_H_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * SBYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank4 = kvu::cluster_ctarank();
    const int pair = (int)(rank4 >> 1), rank = (int)(rank4 & 1);   // pair inside the cluster, CTA inside the pair
    const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
    const int cluster_id = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            kvu::mbar_init(&full[s], 1);
            kvu::mbar_init(&empty[s], 2);      // both pairs' MMA issuers
        }
        for (int a = 0; a < 2; a++) {
            kvu::mbar_init(&tfull[a], 1);
            kvu::mbar_init(&tempty[a], 512);   // 256 epilogue threads of each CTA of the pair
        }
        kvu::fence_barrier_init();
    }
    if (warp == 2) kvu::tmem_alloc2(tmem_slot, 512);
    kvu::tc_fence_before();
    __syncthreads();
    kvu::cluster_sync();
    kvu::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nb = T.n_ptr ? *T.n_ptr : T.n_boards;
    const int NT = T.cout / BN;
    int m_off, M;                                   // this launch's slice of the board tiles: [m_off, m_off + M)
    tower_slice(T, (nb + 3) >> 2, m_off, M);
    const int Ms = (M + 1) >> 1;
    const int m_valid = nb * 64;
    TowerOrder ord;
    ord.init(Ms, NT, T.n_layers, T.chunk_tiles > 1 ? T.chunk_tiles / 2 : T.chunk_tiles);
    const int total = ord.total;
    const uint32_t need = 16u * (uint32_t)NT;

    if (warp == 0) {
        // ---- TMA producer (all four CTAs) ---------------------------------------------------------------------
        int stage = 0;
        uint32_t phase = 0;
        const uint16_t bmask = (uint16_t)((1u << rank) | (1u << (rank + 2)));   // the CTAs that hold this half
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, ms, n_tile;
            ord.decode(t, l, ms, n_tile);
            const int m_tile = m_off + 2 * ms + pair;
            const TowerLayerDev* L = T.layers ? T.layers + l : &T.single;
            const int kb_per_tap = L->kb_per_tap;
            const CUtensorMap* tmH = &maps.h[L->in_buf][L->in_view];
            const int b0 = T.board_base[L->in_view] + m_tile * 4 + rank * 2;
            if (l > 0 && m_tile < m_off + M) {
                tower_wait_tile(T.done + (size_t)(l - 1) * T.m_stride + m_tile, need, lane);
                fence_proxy_async();
            }
            for (int st3 = 0; st3 < 3 * kb_per_tap; st3++) {   // (channel block, dx)
                const int kb = st3 / 3, dxi = st3 - kb * 3;
                kvu::mbar_wait(&empty[stage], phase ^ 1);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * SBYTES;
                    if (rank == 0) kvu::mbar_arrive_expect_tx(&full[stage], 2 * SBYTES);
                    kvu::tma2_load_4d(sa, tmH, &full[stage], kb * BK, dxi - 1, b0, -1);
#pragma unroll
                    for (int dyi = 0; dyi < 3; dyi++)
                        kvu::tma2_load_2d_mc(sa + AH_BYTES + dyi * B2_BYTES + pair * (B2_BYTES / 2), &L->wmap_q, &full[stage],
                                             ((dyi * 3 + dxi) * kb_per_tap + kb) * BK,
                                             n_tile * BN + rank * (BN / 2) + pair * (BN / 4), bmask);
                }
                __syncwarp();
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ---- MMA issuer (the leader CTA of each pair) ------------------------------------------------------------
        constexpr uint32_t idesc = kvu::make_idesc_bf16(2 * BM, BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, ms, n_tile;
            ord.decode(t, l, ms, n_tile);
            const int nst = 3 * (T.layers ? T.layers[l].kb_per_tap : T.single.kb_per_tap);
            kvu::mbar_wait(&tempty[acc], acc_phase ^ 1);
            kvu::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
            for (int s3 = 0; s3 < nst; s3++) {
                kvu::mbar_wait(&full[stage], phase);
                kvu::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = kvu::smem_u32(smem + stage * SBYTES);
#pragma unroll
                    for (int dyi = 0; dyi < 3; dyi++) {
                        const uint64_t adesc = kvu::make_sw128_kmajor_desc(sa + dyi * 2048);
                        const uint64_t bdesc = kvu::make_sw128_kmajor_desc(sa + AH_BYTES + dyi * B2_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; k++)
                            kvu::umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (s3 | dyi | k) != 0);
                    }
                    kvu::umma2_commit_mask(&empty[stage], 0xF);          // the slot is free once BOTH pairs have read it
                    if (s3 == nst - 1) kvu::umma2_commit_mask(&tfull[acc], pair_mask);
                }
                __syncwarp();
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ---- epilogue (all four CTAs, own 128 rows) -----------------------------------------------------------------
        const int q = warp & 3, half = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < total; t += n_clusters) {
            int l, ms, n_tile;
            ord.decode(t, l, ms, n_tile);
            const int m_tile = m_off + 2 * ms + pair;
            const TowerLayerDev* L = T.layers ? T.layers + l : &T.single;
            ConvParams P;
            P.bias = L->bias;
            P.out = T.act[L->out_buf];
            P.residual = L->res_buf >= 0 ? T.act[L->res_buf] : nullptr;
            P.relu = L->relu;
            const int r = q * 32 + lane;   // TMEM lane r = (y, board, x): r / 16, (r / 8) % 2, r % 8
            const int row = m_tile * 2 * BM + rank * BM + ((r >> 3) & 1) * 64 + (r >> 4) * 8 + (r & 7);
            const bool valid = row < m_valid;
            const int colb = n_tile * BN + half * (BN / 2);
            const size_t rbase = (size_t)row * T.cout + (size_t)colb;
            uint4 res[4][4];
            if (P.residual) {
                if (l > 0 && m_tile < m_off + M) tower_wait_tile(T.done + (size_t)(l - 1) * T.m_stride + m_tile, need, lane);
                if (valid) {
                    const uint4* rp = reinterpret_cast<const uint4*>(P.residual + rbase);
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int j = 0; j < 4; j++) res[c][j] = __ldcg(rp + c * 4 + j);
                }
            }
            kvu::mbar_wait(&tfull[acc], acc_phase);
            kvu::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t v[32];
                kvu::tmem_ld_32x32(tmem_base + (uint32_t)acc * BN + half * (BN / 2) + c * 32 + ((uint32_t)(q * 32) << 16), v);
                kvu::tmem_ld_wait();
                if (valid) conv_epilogue_store_pf(P, v, res[c], rbase + c * 32, colb + c * 32);
            }
            kvu::tc_fence_before();
            kvu::mbar_arrive_leader(&tempty[acc]);
            fence_proxy_async();
            __threadfence();
            __syncwarp();
            if (lane == 0 && T.done && m_tile < m_off + M) red_release_add_u32(T.done + (size_t)l * T.m_stride + m_tile, 1u);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    kvu::tc_fence_before();
    __syncthreads();
    kvu::cluster_sync();   // CTAs of the cluster may still be signalling this CTA's barriers / multicasting into it
    if (warp == 2) kvu::tmem_dealloc2(tmem_base, 512);
}

// ---- stem: encode + conv1 + bn1 + relu ------------------------------------------------------------------
// table [9 taps][12 pieces][C1] fp32 (BN scale folded), bias [C1].  One CTA per board.  A warp owns one pixel at a
// time and its lanes own 8 consecutive channels each: table rows are read as coalesced float4 pairs (L1-resident,
// 110 KB), the output row is written as one 512 B (C1 = 256) coalesced burst of 16 B stores.
__global__ void __launch_bounds__(256) stem_kernel(const uint64_t* __restrict__ lines, int n,
                                                   const int* __restrict__ n_ptr,
                                                   const float* __restrict__ table, const float* __restrict__ bias,
                                                   bf16* __restrict__ out, int C1, int pitch) {
    __shared__ int8_t piece[64];
    if (n_ptr) n = *n_ptr;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {   // grid-stride (the pipelined search launches one CTA per SM)
    __syncthreads();                                    // piece[] of the previous board is no longer read
    if (threadIdx.x < 64) {
        int pc = -1;
#pragma unroll
        for (int p = 0; p < 12; p++)
            if ((__ldg(lines + (size_t)b * 16 + p) >> threadIdx.x) & 1) pc = p;
        piece[threadIdx.x] = (int8_t)pc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c0 = lane * 8; c0 < C1; c0 += 256) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
        for (int px = wid; px < 64; px += 8) {
            const int y = px >> 3, x = px & 7;
            float4 a0 = b0, a1 = b1;
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) {
                    const int pc = piece[yy * 8 + xx];
                    if (pc >= 0) {
                        const float4* t = reinterpret_cast<const float4*>(table + ((size_t)tap * 12 + pc) * C1 + c0);
                        const float4 t0 = __ldg(t), t1 = __ldg(t + 1);
                        a0.x += t0.x; a0.y += t0.y; a0.z += t0.z; a0.w += t0.w;
                        a1.x += t1.x; a1.y += t1.y; a1.z += t1.z; a1.w += t1.w;
                    }
                }
            }
            const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f));
            const __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(a0.z, 0.f), fmaxf(a0.w, 0.f));
            const __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f));
            const __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f));
            uint4 o;
            o.x = *reinterpret_cast<const uint32_t*>(&p0);
            o.y = *reinterpret_cast<const uint32_t*>(&p1);
            o.z = *reinterpret_cast<const uint32_t*>(&p2);
            o.w = *reinterpret_cast<const uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(out + ((size_t)b * 64 + px) * pitch + c0) = o;   // pixel pitch of the tower
        }
    }
    }
}

// ---- heads: head_features / value_mlp live in kv_heads.cuh (shared with the search-mode evaluator) ----------
__global__ void __launch_bounds__(256) head_full_kernel(const bf16* __restrict__ act, int n, int C,
                                                        const float* __restrict__ wh, const float* __restrict__ bh,
                                                        const float* __restrict__ wfc, const float* __restrict__ bfc,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        float* __restrict__ policy, float* __restrict__ value) {
    __shared__ float hp[128], hv[64], red[8];
    __shared__ __align__(16) float swh[3 * 512];
    const int b = blockIdx.x;
    if (b >= n) return;
    head_features(act + (size_t)b * 64 * C, C, wh, bh, hp, hv, swh);
    __syncthreads();
    const float v = value_mlp(hv, w1, b1, w2, b2, red);
    if (threadIdx.x == 0 && value) value[b] = v;
    if (policy) {
        for (int o = threadIdx.x; o < 4096; o += blockDim.x) {
            float a = __ldg(bfc + o);
            const float4* wr = reinterpret_cast<const float4*>(wfc + (size_t)o * 128);
#pragma unroll 8
            for (int i = 0; i < 32; i++) {
                const float4 w = __ldg(wr + i);
                a += w.x * hp[4 * i] + w.y * hp[4 * i + 1] + w.z * hp[4 * i + 2] + w.w * hp[4 * i + 3];
            }
            policy[(size_t)b * 4096 + o] = a;
        }
    }
}

// ---- weight preparation (device side) ------------------------------------------------------------------------
// conv weight [Cout][Cin][3][3] fp32 + conv bias + BN(gamma, beta, mean, var) -> [Cout][9][Cin] bf16 + bias fp32
__global__ void fold_conv3x3_kernel(const float* __restrict__ w, const float* __restrict__ cb,
                                    const float* __restrict__ g, const float* __restrict__ be,
                                    const float* __restrict__ mu, const float* __restrict__ var, int cout, int cin,
                                    bf16* __restrict__ wout, float* __restrict__ bout) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t tot = (size_t)cout * 9 * cin;
    if (i < tot) {
        const int ci = (int)(i % cin), tap = (int)((i / cin) % 9), co = (int)(i / ((size_t)cin * 9));
        const float s = g[co] / sqrtf(var[co] + 1e-5f);
        wout[i] = __float2bfloat16(w[((size_t)co * cin + ci) * 9 + tap] * s);
    }
    if (i < (size_t)cout) {
        const float s = g[i] / sqrtf(var[i] + 1e-5f);
        bout[i] = (cb[i] - mu[i]) * s + be[i];
    }
}
// value_fc1.weight [512][64] -> [64][512]
__global__ void transpose_w1_kernel(const float* __restrict__ w, float* __restrict__ wt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 512 * 64) wt[(i & 63) * 512 + (i >> 6)] = w[i];
}
// stem: [C1][12][3][3] -> table [9][12][C1] fp32
__global__ void fold_stem_kernel(const float* __restrict__ w, const float* __restrict__ cb, const float* __restrict__ g,
                                 const float* __restrict__ be, const float* __restrict__ mu,
                                 const float* __restrict__ var, int c1, float* __restrict__ table,
                                 float* __restrict__ bout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * 12 * c1) {
        const int co = i % c1, pc = (i / c1) % 12, tap = i / (c1 * 12);
        const float s = g[co] / sqrtf(var[co] + 1e-5f);
        table[i] = w[((size_t)co * 12 + pc) * 9 + tap] * s;
    }
    if (i < c1) {
        const float s = g[i] / sqrtf(var[i] + 1e-5f);
        bout[i] = (cb[i] - mu[i]) * s + be[i];
    }
}
// 1x1 head convs: rows [r0, r0+rows) of wh [3][C]
__global__ void fold_head_kernel(const float* __restrict__ w, const float* __restrict__ cb, const float* __restrict__ g,
                                 const float* __restrict__ be, const float* __restrict__ mu,
                                 const float* __restrict__ var, int rows, int C, float* __restrict__ wh,
                                 float* __restrict__ bh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * C) {
        const int r = i / C;
        const float s = g[r] / sqrtf(var[r] + 1e-5f);
        wh[i] = w[i] * s;
    }
    if (i < rows) {
        const float s = g[i] / sqrtf(var[i] + 1e-5f);
        bh[i] = (cb[i] - mu[i]) * s + be[i];
    }
}

// planes [n][12][8][8] float (one-hot) -> board lines (bitboards only; meta = 0)
__global__ void planes_to_lines_kernel(const float* __restrict__ planes, int n, uint64_t* __restrict__ lines,
                                       int* __restrict__ not_onehot) {
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;   // 12 warps, one per plane
    if (b >= n || wid >= 12) return;
    const float* p = planes + ((size_t)b * 12 + wid) * 64;
    const float a = p[lane], c = p[32 + lane];
    if ((a != 0.f && a != 1.f) || (c != 0.f && c != 1.f)) atomicExch(not_onehot, 1);
    const uint32_t lo = __ballot_sync(0xffffffffu, a != 0.f), hi = __ballot_sync(0xffffffffu, c != 0.f);
    if (lane == 0) lines[(size_t)b * 16 + wid] = (uint64_t)lo | ((uint64_t)hi << 32);
    if (wid == 0 && lane < 4) lines[(size_t)b * 16 + 12 + lane] = 0;
}

}  // namespace kvn

using namespace kvn;

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// pitch = elements between consecutive pixels (0: dense, = C)
int kv_make_act_map(kv_ctx* ctx, CUtensorMap* m, void* base, int C, int boards, int box_boards, int pitch) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled unavailable");
    if (pitch <= 0) pitch = C;
    cuuint64_t dims[4] = {(cuuint64_t)C, 8, 8, (cuuint64_t)boards};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 16, (cuuint64_t)pitch * 128};
    cuuint32_t box[4] = {64, 8, 8, (cuuint32_t)box_boards};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled(activations) failed");
    return 0;
}
// halo variant: dims (channel, x, board, y) so that a box {64, 8, 2, 10} lands as rows [y][board][x]
static int make_act_map_h(kv_ctx* ctx, CUtensorMap* m, void* base, int C, int boards, int pitch) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[4] = {(cuuint64_t)C, 8, (cuuint64_t)boards, 8};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 128, (cuuint64_t)pitch * 16};
    cuuint32_t box[4] = {64, 8, 2, 10};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled(activations, halo layout) failed");
    return 0;
}
static int make_act_map(kv_ctx* ctx, CUtensorMap* m, void* base, int C, int boards, int pitch) {
    return kv_make_act_map(ctx, m, base, C, boards, 2, pitch);
}
static int make_w_map(kv_ctx* ctx, CUtensorMap* m, void* base, int cout, int K, int box_rows = 256) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kv_fail_msg(ctx, "cuTensorMapEncodeTiled(weights) failed");
    return 0;
}

static size_t net_blob_floats(const kv_net* n) {
    auto conv_bn = [](size_t cout, size_t cin, size_t k) { return cout * cin * k * k + cout + 4 * cout; };
    size_t t = conv_bn(n->C1, 12, 3);
    if (n->has_conv2) t += conv_bn(n->C, n->C1, 3);
    t += (size_t)n->blocks * 2 * conv_bn(n->C, n->C, 3);
    t += conv_bn(2, n->C, 1) + 4096 * 128 + 4096;
    t += conv_bn(1, n->C, 1) + 512 * 64 + 512 + 512 + 1;
    return t;
}

// buffer rotation of the tower: conv2 (0 -> 1), then per residual block x -> t -> o (+ x), x = o
struct TowerStep {
    int in, out, res;
};
static std::vector<TowerStep> tower_plan(const kv_net* n) {
    std::vector<TowerStep> p;
    int x = 0;
    if (n->has_conv2) {
        p.push_back({0, 1, -1});
        x = 1;
    }
    for (int b = 0; b < n->blocks; b++) {
        const int t = (x + 1) % 3, o = (x + 2) % 3;
        p.push_back({x, t, -1});
        p.push_back({t, o, x});
        x = o;
    }
    return p;
}

void kv_net_destroy(kv_ctx* ctx) {
    kv_net* n = ctx->net;
    if (!n) return;
    for (int i = 0; i < 3; i++)
        if (n->act[i]) cudaFree(n->act[i]);
    // the folded weights (tower bf16 + biases, stem table, heads) are slices of one arena
    void* ptrs[] = {n->d_folded, n->d_blob, n->d_flag, n->d_lines_tmp, n->d_layers, n->d_done[0], n->d_done[1]};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (n->side) cudaStreamDestroy(n->side);
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    delete n;
    ctx->net = nullptr;
}

extern "C" {

int kv_net_create(kv_ctx* ctx, int stem_channels, int tower_channels, int n_blocks, int has_conv2, int max_boards) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stem_channels % 64 || tower_channels % 256 || stem_channels > 512 || tower_channels > 512 || n_blocks < 0 ||
        max_boards < 1)
        return kv_fail_msg(ctx, "kv_net_create: channels must be multiples of 64 (stem) / 256 (tower), <= 512");
    if (stem_channels > tower_channels)
        return kv_fail_msg(ctx, "kv_net_create: the stem must not be wider than the tower");
    if (!has_conv2 && stem_channels != tower_channels)
        return kv_fail_msg(ctx, "kv_net_create: without conv2 the stem must produce the tower width");
    kv_net_destroy(ctx);
    kv_net* n = new kv_net();
    ctx->net = n;
    n->C1 = stem_channels;
    n->C = tower_channels;
    n->blocks = n_blocks;
    n->has_conv2 = has_conv2 != 0;
    n->cap = (max_boards + 1) & ~1;
    const int cmax = n->C > n->C1 ? n->C : n->C1;
    for (int i = 0; i < 3; i++) {
        KV_CUDA(ctx, cudaMalloc(&n->act[i], (size_t)n->cap * 64 * cmax * sizeof(bf16)));
        KV_CUDA(ctx, cudaMemset(n->act[i], 0, (size_t)n->cap * 64 * cmax * sizeof(bf16)));
        // every view has the pixel pitch of the widest layer (cmax channels): board b lives at byte b * 64 * cmax * 2 in
        // every view, so a narrower tensor (the stem output) aliases only the SAME board of a wider one — the
        // dependency-scheduled tower relies on that (a tile's output may only overwrite data of its own boards)
        if (int rc = make_act_map(ctx, &n->map_act[i][0], n->act[i], n->C1, n->cap, cmax)) return rc;
        if (int rc = make_act_map(ctx, &n->map_act[i][1], n->act[i], n->C, n->cap, cmax)) return rc;
        n->halo_ok = n->halo_ok && make_act_map_h(ctx, &n->map_act_h[i][0], n->act[i], n->C1, n->cap, cmax) == 0 &&
                     make_act_map_h(ctx, &n->map_act_h[i][1], n->act[i], n->C, n->cap, cmax) == 0;
    }
    const int nconv = (n->has_conv2 ? 1 : 0) + 2 * n->blocks;
    n->convs.resize(nconv);
    // The FOLDED weights — what the kernels read: tower weights bf16 with the BatchNorm scale folded in, fp32 biases,
    // the stem table, the heads — live in ONE arena (256 B aligned slices), so a generation's weights can travel between
    // GPUs as a single 52 MB NCCL broadcast of this arena (half the fp32 state_dict blob) with no fold on the receivers.
    size_t off = 0;
    auto slice = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    std::vector<size_t> o_w(nconv), o_b(nconv);
    for (int l = 0; l < nconv; l++) {
        kv_conv& L = n->convs[l];
        L.cin = (n->has_conv2 && l == 0) ? n->C1 : n->C;
        L.cout = n->C;
        o_w[l] = slice((size_t)L.cout * 9 * L.cin * sizeof(bf16));
        o_b[l] = slice((size_t)L.cout * sizeof(float));
    }
    const size_t o_st = slice((size_t)9 * 12 * n->C1 * sizeof(float)), o_sb = slice((size_t)n->C1 * sizeof(float));
    const size_t o_wh = slice((size_t)3 * n->C * sizeof(float)), o_bh = slice(4 * sizeof(float));
    const size_t o_wfc = slice((size_t)4096 * 128 * sizeof(float)), o_bfc = slice(4096 * sizeof(float));
    const size_t o_w1 = slice(512 * 64 * sizeof(float)), o_b1 = slice(512 * sizeof(float));
    const size_t o_w2 = slice(512 * sizeof(float)), o_b2 = slice(4 * sizeof(float));
    n->folded_bytes = off;
    KV_CUDA(ctx, cudaMalloc(&n->d_folded, n->folded_bytes));
    KV_CUDA(ctx, cudaMemset(n->d_folded, 0, n->folded_bytes));
    char* fb = static_cast<char*>(n->d_folded);
    for (int l = 0; l < nconv; l++) {
        kv_conv& L = n->convs[l];
        L.w = reinterpret_cast<bf16*>(fb + o_w[l]);
        L.b = reinterpret_cast<float*>(fb + o_b[l]);
        if (int rc = make_w_map(ctx, &L.map, L.w, L.cout, 9 * L.cin)) return rc;
        if (int rc = make_w_map(ctx, &L.map_half, L.w, L.cout, 9 * L.cin, 128)) return rc;
        if (int rc = make_w_map(ctx, &L.map_q, L.w, L.cout, 9 * L.cin, 64)) return rc;
    }
    n->stem_table = reinterpret_cast<float*>(fb + o_st);
    n->stem_bias = reinterpret_cast<float*>(fb + o_sb);
    n->wh = reinterpret_cast<float*>(fb + o_wh);
    n->bh = reinterpret_cast<float*>(fb + o_bh);
    n->wfc = reinterpret_cast<float*>(fb + o_wfc);
    n->bfc = reinterpret_cast<float*>(fb + o_bfc);
    n->w1 = reinterpret_cast<float*>(fb + o_w1);
    n->b1 = reinterpret_cast<float*>(fb + o_b1);
    n->w2 = reinterpret_cast<float*>(fb + o_w2);
    n->b2 = reinterpret_cast<float*>(fb + o_b2);
    KV_CUDA(ctx, cudaMalloc(&n->d_flag, 4 * sizeof(int)));
    KV_CUDA(ctx, cudaMalloc(&n->d_lines_tmp, (size_t)n->cap * 128));
    n->blob_floats = net_blob_floats(n);
    KV_CUDA(ctx, cudaMalloc(&n->d_blob, n->blob_floats * sizeof(float)));
    KV_CUDA(ctx, cudaFuncSetAttribute(conv3x3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    if (const char* e = getenv("KV_CONV_CTA_GROUP")) n->conv_mode = atoi(e) == 1 ? 1 : 2;
    // the whole-tower launch: per-layer table (weight map, bias, buffer rotation) and the tile-completion counters
    KV_CUDA(ctx, cudaFuncSetAttribute(tower_umma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV2_SMEM));
    KV_CUDA(ctx, cudaFuncSetAttribute(tower_umma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_H_SMEM));
    if (nconv > 0) {
        std::vector<TowerLayerDev> tab(nconv);
        std::vector<TowerStep> plan = tower_plan(n);
        for (int l = 0; l < nconv; l++) {
            memset(static_cast<void*>(&tab[l]), 0, sizeof(TowerLayerDev));
            tab[l].wmap = n->convs[l].map_half;
            tab[l].wmap_q = n->convs[l].map_q;
            tab[l].bias = n->convs[l].b;
            tab[l].in_buf = plan[l].in;
            tab[l].in_view = (n->convs[l].cin == n->C1 && n->convs[l].cin != n->C) ? 0 : 1;
            tab[l].out_buf = plan[l].out;
            tab[l].res_buf = plan[l].res;
            tab[l].kb_per_tap = n->convs[l].cin / BK;
            tab[l].relu = 1;
        }
        KV_CUDA(ctx, cudaMalloc(&n->d_layers, nconv * sizeof(TowerLayerDev)));
        KV_CUDA(ctx, cudaMemcpy(n->d_layers, tab.data(), nconv * sizeof(TowerLayerDev), cudaMemcpyHostToDevice));
        n->m_stride = (n->cap + 3) / 4;
        for (int i = 0; i < 2; i++) KV_CUDA(ctx, cudaMalloc(&n->d_done[i], (size_t)nconv * n->m_stride * sizeof(uint32_t)));
    }
    {   // 4-CTA clusters: how many are co-resident (the dependency schedule needs every cluster of the grid resident)
        cudaError_t e = cudaFuncSetAttribute(tower_umma4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_H_SMEM);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(ctx->sm_count / 4 * 4));
        cfg.blockDim = dim3(CONV2_THREADS);
        cfg.dynamicSmemBytes = TOWER_H_SMEM;
        cudaLaunchAttribute at;
        memset(&at, 0, sizeof(at));
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 4;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int nc = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&nc, tower_umma4_kernel<0>, &cfg);
        n->clusters4 = (e == cudaSuccess && n->halo_ok) ? nc : 0;
        cudaGetLastError();
    }
    if (const char* e = getenv("KV_TOWER_FUSED")) n->tower_fused = atoi(e);
    if (n->tower_fused < 0 || n->tower_fused > 4 || (n->tower_fused == 2 && !n->halo_ok) ||
        (n->tower_fused >= 3 && n->clusters4 < 8))
        n->tower_fused = 1;
    // depth-first chunks: one round of the CTA pairs per layer and channel tile, 57-85 MB of activations per chunk at
    // 512 channels (more tiles for a narrower tower: same bytes)
    n->tower_chunk = (ctx->sm_count / 2) * (512 / n->C);
    if (const char* e = getenv("KV_TOWER_CHUNK")) n->tower_chunk = atoi(e);
    return 0;
}

// 1 (default) = the tower convolutions of a forward pass run as ONE dependency-scheduled launch (tower_umma2_kernel),
// 0 = one launch per layer.  Same tiles, same arithmetic: bit-identical outputs.
int kv_net_set_tower_fused(kv_ctx* ctx, int on) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_set_tower_fused: no net");
    if (on < 0 || on > 4) return kv_fail_msg(ctx, "kv_net_set_tower_fused: 0 .. 4");
    if (on >= 3 && ctx->net->clusters4 < 8) return kv_fail_msg(ctx, "kv_net_set_tower_fused: 4-CTA clusters unavailable");
    if (on == 2 && !ctx->net->halo_ok) return kv_fail_msg(ctx, "kv_net_set_tower_fused: halo tensor maps unavailable");
    ctx->net->tower_fused = on;
    return 0;
}

// 1 = one CTA per tile (cta_group::1), 2 = CTA pairs (cta_group::2, default).  Same results bit for bit.
int kv_net_set_conv_mode(kv_ctx* ctx, int cta_group) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_set_conv_mode: no net");
    if (cta_group != 1 && cta_group != 2) return kv_fail_msg(ctx, "kv_net_set_conv_mode: 1 or 2");
    ctx->net->conv_mode = cta_group;
    return 0;
}

// co-resident 4-CTA clusters of the multicast tower kernel (0: kv_net_set_tower_fused(3) is unavailable)
int kv_net_tower_clusters4(kv_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->clusters4 : 0; }

uint64_t kv_net_blob_floats(kv_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->blob_floats : 0; }
void* kv_net_blob_device_ptr(kv_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->d_blob : nullptr; }

// Fold the fp32 state_dict blob that sits in the net's device staging buffer (kv_net_blob_device_ptr).
int kv_mcts_cache_clear(kv_ctx* ctx, void* stream);

int kv_net_commit_weights(kv_ctx* ctx, void* stream) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_commit_weights: no net");
    kv_net* n = ctx->net;
    cudaStream_t st = (cudaStream_t)stream;
    const float* p = n->d_blob;
    auto take = [&](size_t k) {
        const float* r = p;
        p += k;
        return r;
    };
    {   // conv1 + bn1
        const float* w = take((size_t)n->C1 * 12 * 9);
        const float* cb = take(n->C1);
        const float *g = take(n->C1), *be = take(n->C1), *mu = take(n->C1), *var = take(n->C1);
        const int tot = 9 * 12 * n->C1;
        fold_stem_kernel<<<(tot + 255) / 256, 256, 0, st>>>(w, cb, g, be, mu, var, n->C1, n->stem_table, n->stem_bias);
        KV_LAUNCH_CHECK(ctx);
    }
    for (auto& L : n->convs) {
        const float* w = take((size_t)L.cout * L.cin * 9);
        const float* cb = take(L.cout);
        const float *g = take(L.cout), *be = take(L.cout), *mu = take(L.cout), *var = take(L.cout);
        const size_t tot = (size_t)L.cout * 9 * L.cin;
        fold_conv3x3_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(w, cb, g, be, mu, var, L.cout, L.cin, L.w, L.b);
        KV_LAUNCH_CHECK(ctx);
    }
    {   // policy head
        const float* w = take((size_t)2 * n->C);
        const float* cb = take(2);
        const float *g = take(2), *be = take(2), *mu = take(2), *var = take(2);
        fold_head_kernel<<<(2 * n->C + 255) / 256, 256, 0, st>>>(w, cb, g, be, mu, var, 2, n->C, n->wh, n->bh);
        KV_LAUNCH_CHECK(ctx);
        KV_CUDA(ctx, cudaMemcpyAsync(n->wfc, take((size_t)4096 * 128), (size_t)4096 * 128 * 4, cudaMemcpyDeviceToDevice, st));
        KV_CUDA(ctx, cudaMemcpyAsync(n->bfc, take(4096), 4096 * 4, cudaMemcpyDeviceToDevice, st));
    }
    {   // value head
        const float* w = take((size_t)n->C);
        const float* cb = take(1);
        const float *g = take(1), *be = take(1), *mu = take(1), *var = take(1);
        fold_head_kernel<<<(n->C + 255) / 256, 256, 0, st>>>(w, cb, g, be, mu, var, 1, n->C, n->wh + 2 * n->C, n->bh + 2);
        KV_LAUNCH_CHECK(ctx);
        // value_fc1 is kept TRANSPOSED ([64 inputs][512 units]): consecutive threads own consecutive units, so every
        // weight load of the value MLP is one coalesced 128 B request instead of 32 scattered 16 B ones
        transpose_w1_kernel<<<(512 * 64 + 255) / 256, 256, 0, st>>>(take(512 * 64), n->w1);
        KV_LAUNCH_CHECK(ctx);
        KV_CUDA(ctx, cudaMemcpyAsync(n->b1, take(512), 512 * 4, cudaMemcpyDeviceToDevice, st));
        KV_CUDA(ctx, cudaMemcpyAsync(n->w2, take(512), 512 * 4, cudaMemcpyDeviceToDevice, st));
        KV_CUDA(ctx, cudaMemcpyAsync(n->b2, take(1), 4, cudaMemcpyDeviceToDevice, st));
    }
    if ((size_t)(p - n->d_blob) != n->blob_floats) return kv_fail_msg(ctx, "kv_net_commit_weights: blob size mismatch");
    n->loaded = true;
    return kv_mcts_cache_clear(ctx, stream);   // cached features belong to the old weights
}

// The folded weights as one device buffer (see kv_net_create): after kv_net_commit_weights on the source rank, broadcast
// these bytes into every other rank's arena and call kv_net_adopt_folded there — no fp32 blob, no fold on the receivers.
uint64_t kv_net_folded_bytes(kv_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->folded_bytes : 0; }
void* kv_net_folded_device_ptr(kv_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->d_folded : nullptr; }
int kv_net_adopt_folded(kv_ctx* ctx, void* stream) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_adopt_folded: no net");
    ctx->net->loaded = true;
    return kv_mcts_cache_clear(ctx, stream);   // cached features belong to the old weights
}

int kv_net_load(kv_ctx* ctx, const float* h_blob, uint64_t n_floats) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_load: no net");
    if (n_floats != ctx->net->blob_floats) return kv_fail_msg(ctx, "kv_net_load: blob has the wrong number of floats");
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    KV_CUDA(ctx, cudaMemcpy(ctx->net->d_blob, h_blob, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    if (int rc = kv_net_commit_weights(ctx, nullptr)) return rc;
    KV_CUDA(ctx, cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"

// One 3x3 convolution on caller-owned NHWC bf16 tensors (the training path, kv_train.cu): the tower's CTA-pair
// implicit-GEMM kernel with tensor maps built for this call.  y = [relu](conv(x, w) + bias [+ residual]).
int kv_conv_launch(kv_ctx* ctx, const bf16* x, const bf16* w_packed, const float* bias, const bf16* residual, bf16* y,
                   int n, int cin, int cout, int relu, cudaStream_t st) {
    if (cin % 64 || cout % 256 || cin < 64) return kv_fail_msg(ctx, "conv3x3: cin must be a multiple of 64, cout of 256");
    CUtensorMap wmap;
    if (int rc = make_w_map(ctx, &wmap, const_cast<bf16*>(w_packed), cout, 9 * cin, 128)) return rc;
    static const bool halo = [] {
        const char* e = getenv("KV_CONV_HALO");
        return !(e && atoi(e) == 0);
    }();
    // a one-layer launch of the tower kernel: x is buffer 0, y buffer 1, the residual buffer 2
    TowerActMaps maps;              // only [0][1] of the variant's map set is read
    memset(static_cast<void*>(&maps), 0, sizeof(maps));
    if (halo) {
        if (int rc = make_act_map_h(ctx, &maps.h[0][1], const_cast<bf16*>(x), cin, n, cin)) return rc;
    } else {
        if (int rc = kv_make_act_map(ctx, &maps.m[0][1], const_cast<bf16*>(x), cin, n, 2)) return rc;
    }
    TowerArgs T;
    memset(static_cast<void*>(&T), 0, sizeof(T));
    T.act[0] = const_cast<bf16*>(x);
    T.act[1] = y;
    T.act[2] = const_cast<bf16*>(residual);
    T.n_layers = 1;
    T.n_boards = n;
    T.cout = cout;
    T.single.wmap = wmap;
    T.single.bias = bias;
    T.single.in_buf = 0;
    T.single.in_view = 1;
    T.single.out_buf = 1;
    T.single.res_buf = residual ? 2 : -1;
    T.single.kb_per_tap = cin / BK;
    T.single.relu = relu;
    if (!ctx->conv_attr_done) {
        KV_CUDA(ctx, cudaFuncSetAttribute(tower_umma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_H_SMEM));
        KV_CUDA(ctx, cudaFuncSetAttribute(tower_umma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV2_SMEM));
        ctx->conv_attr_done = true;
    }
    const int pairs = ctx->sm_count / 2;
    const int total = ((n + 3) / 4) * (cout / BN);
    const int grid = 2 * (total < pairs ? total : pairs);
    KvTimed t_(ctx, KVK_NET_CONV, st);
    if (halo) tower_umma2_kernel<true><<<grid, CONV2_THREADS, TOWER_H_SMEM, st>>>(maps, T);
    else tower_umma2_kernel<false><<<grid, CONV2_THREADS, CONV2_SMEM, st>>>(maps, T);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

// Runs stem + tower for n boards; returns the buffer index holding the final activations.
int kv_net_tower(kv_ctx* ctx, const uint64_t* d_lines, int n, cudaStream_t st, int* final_buf, int max_convs,
                 const int* n_ptr, int board_base, cudaStream_t conv_stream, cudaEvent_t handoff, int stem_grid) {
    kv_net* net = ctx->net;
    if (!net || !net->loaded) return kv_fail_msg(ctx, "net: weights not loaded");
    if (board_base < 0 || board_base + n > net->cap)
        return kv_fail_msg(ctx, "net: batch exceeds max_boards given to kv_net_create");
    const int cmax = net->C;   // C1 <= C (kv_net_create): every activation row has the tower's pixel pitch
    const size_t off0 = (size_t)board_base * 64 * cmax;   // element offset of this launch in every activation buffer
    {
        KvTimed t_(ctx, KVK_NET_STEM, st);
        stem_kernel<<<(stem_grid > 0 && stem_grid < n) ? stem_grid : n, 256, 0, st>>>(d_lines, n, n_ptr, net->stem_table,
                                                                                      net->stem_bias, net->act[0] + off0, net->C1, cmax);
    }
    KV_LAUNCH_CHECK(ctx);
    // pipelined search: the tensor-core kernels of both game groups run on one (high-priority) stream, in issue order;
    // everything up to here ran on the group's own stream and overlaps the other group's tower
    cudaStream_t cs = st;
    if (conv_stream && handoff) {
        KV_CUDA(ctx, cudaEventRecord(handoff, st));
        KV_CUDA(ctx, cudaStreamWaitEvent(conv_stream, handoff, 0));
        cs = conv_stream;
    }
    int x = 0;   // buffer holding the current block input
    auto conv = [&](int layer, int in, int out, int res, int relu) -> int {
        kv_conv& L = net->convs[layer];
        ConvParams P;
        P.bias = L.b;
        P.residual = res >= 0 ? net->act[res] + off0 : nullptr;
        P.out = net->act[out] + off0;
        P.board_base = board_base;
        P.m_tiles = (n + 1) / 2;
        P.n_tiles = L.cout / BN;
        P.kb_per_tap = L.cin / BK;
        P.cout = L.cout;
        P.m_valid = n * 64;
        P.relu = relu;
        P.n_ptr = n_ptr;
        const CUtensorMap& amap = net->map_act[in][L.cin == net->C1 && L.cin != net->C ? 0 : 1];
        if (net->conv_mode == 2) {   // one-layer launch of the CTA-pair kernel (9-fetch variant)
            TowerActMaps maps;
            for (int i = 0; i < 3; i++)
                for (int v = 0; v < 2; v++) {
                    maps.m[i][v] = net->map_act[i][v];
                    maps.h[i][v] = net->map_act_h[i][v];
                }
            TowerArgs T;
            memset(static_cast<void*>(&T), 0, sizeof(T));
            for (int i = 0; i < 3; i++) T.act[i] = net->act[i] + off0;
            T.n_ptr = n_ptr;
            T.n_layers = 1;
            T.n_boards = n;
            T.cout = L.cout;
            T.board_base[0] = T.board_base[1] = board_base;
            T.single.wmap = L.map_half;
            T.single.bias = L.b;
            T.single.in_buf = in;
            T.single.in_view = (L.cin == net->C1 && L.cin != net->C) ? 0 : 1;
            T.single.out_buf = out;
            T.single.res_buf = res;
            T.single.kb_per_tap = L.cin / BK;
            T.single.relu = relu;
            const int total = ((n + 3) / 4) * P.n_tiles;
            const int pairs = ctx->sm_count / 2;
            const int grid = 2 * (total < pairs ? total : pairs);
            KvTimed t_(ctx, KVK_NET_CONV, cs);
            tower_umma2_kernel<false><<<grid, CONV2_THREADS, CONV2_SMEM, cs>>>(maps, T);
        } else {
            const int total = P.m_tiles * P.n_tiles;
            const int grid = total < ctx->sm_count ? total : ctx->sm_count;
            KvTimed t_(ctx, KVK_NET_CONV, cs);
            conv3x3_umma_kernel<<<grid, CONV_THREADS, CONV_SMEM, cs>>>(amap, L.map, P);
        }
        KV_LAUNCH_CHECK(ctx);
        return 0;
    };
    int layer = 0;
    *final_buf = 0;
    if (max_convs == 0) return 0;
    if (net->conv_mode == 2 && net->tower_fused && !net->convs.empty()) {
        const std::vector<TowerStep> plan = tower_plan(net);
        const int nl = (max_convs < 0 || max_convs > (int)plan.size()) ? (int)plan.size() : max_convs;
        *final_buf = plan[nl - 1].out;
        uint32_t* done = net->d_done[board_base ? 1 : 0];
        KV_CUDA(ctx, cudaMemsetAsync(done, 0, (size_t)nl * net->m_stride * sizeof(uint32_t), cs));
        TowerActMaps maps;
        for (int i = 0; i < 3; i++)
            for (int v = 0; v < 2; v++) {
                maps.m[i][v] = net->map_act[i][v];
                maps.h[i][v] = net->map_act_h[i][v];
            }
        TowerArgs T;
        memset(static_cast<void*>(&T), 0, sizeof(T));
        for (int i = 0; i < 3; i++) T.act[i] = net->act[i] + off0;
        T.layers = reinterpret_cast<const TowerLayerDev*>(net->d_layers);
        T.done = done;
        T.n_ptr = n_ptr;
        T.n_layers = nl;
        T.m_stride = net->m_stride;
        T.n_boards = n;
        T.cout = net->C;
        T.board_base[0] = T.board_base[1] = board_base;
        T.chunk_tiles = net->tower_chunk;
        const int total = nl * ((n + 3) / 4) * (net->C / BN);
        const int pairs = ctx->sm_count / 2;
        const int grid = 2 * (total < pairs ? total : pairs);
        {
            KvTimed t_(ctx, KVK_NET_CONV, cs);
            if (net->tower_fused == 4 && (n + 3) / 4 >= 4 * net->clusters4) {
                // hybrid: the 4-CTA clusters cannot cover every SM (33 clusters = 132 of 148 SMs on B200: GPCs whose SM
                // count is not a multiple of four), so the multicast kernel takes the share of the board tiles that matches
                // its SMs and the CTA-pair kernel runs the rest on the remaining TPCs, concurrently on a second stream
                if (!net->side) {
                    KV_CUDA(ctx, cudaStreamCreateWithFlags(&net->side, cudaStreamNonBlocking));
                    KV_CUDA(ctx, cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
                    KV_CUDA(ctx, cudaEventCreateWithFlags(&net->ev_join, cudaEventDisableTiming));
                }
                const int sm4 = 4 * net->clusters4, rest = ctx->sm_count - sm4;
                T.split_num = sm4;
                T.split_den = sm4 + rest;
                KV_CUDA(ctx, cudaEventRecord(net->ev_fork, cs));
                KV_CUDA(ctx, cudaStreamWaitEvent(net->side, net->ev_fork, 0));
                T.part = 1;
                tower_umma4_kernel<0><<<sm4, CONV2_THREADS, TOWER_H_SMEM, cs>>>(maps, T);
                T.part = 2;
                tower_umma2_kernel<true><<<rest & ~1, CONV2_THREADS, TOWER_H_SMEM, net->side>>>(maps, T);
                KV_CUDA(ctx, cudaEventRecord(net->ev_join, net->side));
                KV_CUDA(ctx, cudaStreamWaitEvent(cs, net->ev_join, 0));
            } else if (net->tower_fused == 3 || net->tower_fused == 4) {
                const int total4 = nl * (((n + 3) / 4 + 1) / 2) * (net->C / BN);
                const int nc = total4 < net->clusters4 ? total4 : net->clusters4;
                tower_umma4_kernel<0><<<4 * nc, CONV2_THREADS, TOWER_H_SMEM, cs>>>(maps, T);
            } else if (net->tower_fused == 2) tower_umma2_kernel<true><<<grid, CONV2_THREADS, TOWER_H_SMEM, cs>>>(maps, T);
            else tower_umma2_kernel<false><<<grid, CONV2_THREADS, CONV2_SMEM, cs>>>(maps, T);
        }
        KV_LAUNCH_CHECK(ctx);
        return 0;
    }
    if (net->has_conv2) {
        if (int rc = conv(layer++, 0, 1, -1, 1)) return rc;
        x = 1;
        *final_buf = x;
        if (layer == max_convs) return 0;
    }
    for (int b = 0; b < net->blocks; b++) {
        const int t = (x + 1) % 3, o = (x + 2) % 3;
        if (int rc = conv(layer++, x, t, -1, 1)) return rc;
        if (layer == max_convs) {   // debug stop inside a block: expose the intermediate
            *final_buf = t;
            return 0;
        }
        if (int rc = conv(layer++, t, o, x, 1)) return rc;
        x = o;
        *final_buf = x;
        if (layer == max_convs) return 0;
    }
    return 0;
}

extern "C" {

int kv_net_forward(kv_ctx* ctx, const uint64_t* d_lines, int n, float* d_policy, float* d_value, void* stream) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int fb = 0;
    if (int rc = kv_net_tower(ctx, d_lines, n, st, &fb, -1)) return rc;
    kv_net* net = ctx->net;
    {
        KvTimed t_(ctx, KVK_NET_HEAD, st);
        head_full_kernel<<<n, 256, 0, st>>>(net->act[fb], n, net->C, net->wh, net->bh, net->wfc, net->bfc, net->w1,
                                            net->b1, net->w2, net->b2, d_policy, d_value);
    }
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

// Debug/test hook: run the stem and the first n_convs tower convolutions, copy the bf16 NHWC activations out.
int kv_net_forward_partial(kv_ctx* ctx, const uint64_t* d_lines, int n, int n_convs, void* d_act_out, int* channels) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_forward_partial: no net");
    int fb = 0;
    if (int rc = kv_net_tower(ctx, d_lines, n, nullptr, &fb, n_convs)) return rc;
    const int C = n_convs == 0 ? ctx->net->C1 : ctx->net->C;
    if (channels) *channels = C;
    // rows of C channels at the tower's pixel pitch -> dense [n*64][C]
    KV_CUDA(ctx, cudaMemcpy2DAsync(d_act_out, (size_t)C * 2, ctx->net->act[fb], (size_t)ctx->net->C * 2, (size_t)C * 2,
                                   (size_t)n * 64, cudaMemcpyDeviceToDevice, nullptr));
    KV_CUDA(ctx, cudaStreamSynchronize(nullptr));
    return 0;
}

// ChessNet.forward drop-in: fp32 one-hot planes in (the reference's input format), logits + value out.
int kv_net_forward_planes(kv_ctx* ctx, const float* d_planes, int n, float* d_policy, float* d_value, void* stream) {
    if (!ctx || !ctx->net) return kv_fail_msg(ctx, "kv_net_forward_planes: no net");
    if (n <= 0) return 0;
    kv_net* net = ctx->net;
    if (n > net->cap) return kv_fail_msg(ctx, "net: batch exceeds max_boards given to kv_net_create");
    cudaStream_t st = (cudaStream_t)stream;
    KV_CUDA(ctx, cudaMemsetAsync(net->d_flag, 0, sizeof(int), st));
    planes_to_lines_kernel<<<n, 384, 0, st>>>(d_planes, n, net->d_lines_tmp, net->d_flag);
    KV_LAUNCH_CHECK(ctx);
    int flag = 0;
    KV_CUDA(ctx, cudaMemcpyAsync(&flag, net->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    KV_CUDA(ctx, cudaStreamSynchronize(st));
    if (flag) return kv_fail_msg(ctx, "kv_net_forward_planes: input planes are not 0/1 (encode_board output expected)");
    return kv_net_forward(ctx, net->d_lines_tmp, n, d_policy, d_value, stream);
}

}  // extern "C"
