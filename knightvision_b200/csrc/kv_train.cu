// kv_train.cu — the training-side 3x3 convolution operators of the policy/value tower on sm_100a tensor cores.
//
// The reference trains ai/model.py with PyTorch autocast + cuDNN (scripts/train.py:126-196).  Per tower layer that is
// three GEMMs; here each runs on tcgen05 over caller-owned NHWC bf16 tensors [boards][8][8][C]:
//   fprop  Y = conv(X, W) + b                          the tower's CTA-pair implicit-GEMM kernel (kv_net.cu)
//   dgrad  dX = conv(dY, W'),  W'[ci][tap][co] = W[co][8 - tap][ci]       the same kernel on re-packed weights
//   wgrad  dW[co][tap][ci] = sum over pixels of dY[p][co] * X[p + tap][ci]      conv3x3_wgrad_kernel below
// wgrad is a GEMM whose reduction dimension is the PIXEL index, which is the slow dimension of NHWC.  Instead of
// transposing the activations it uses the MN-major operand mode of tcgen05: a TMA box {64 channels, 8, 8, 1 board}
// lands in shared memory as 64 pixel rows x 128 B (SWIZZLE_128B), which is the canonical MN-major layout with
// 8-row K groups 1024 B apart (stride byte offset) and 64-channel MN groups one box (8 KB) apart (leading byte
// offset).  The tap shift of X is the TMA start coordinate (dx, dy) with out-of-board pixels zero-filled, exactly as
// in the forward kernel.  One CTA owns one (128 output channels) x (256 input channels) x tap tile and a slice of
// the boards (split-K so that ~148 CTAs run); fp32 partials go to a workspace and a second kernel adds them in a
// fixed order (deterministic) while writing the [Cout][Cin][3][3] layout of the PyTorch parameter.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>

#include "kv_internal.h"
#include "kv_net.h"
#include "kv_umma.cuh"

using bf16 = __nv_bfloat16;

namespace kvt {

constexpr int WG_STAGES = 4;
constexpr int WG_BOX = 64 * 64 * 2;              // one TMA box: 64 pixels x 64 channels bf16 = 8 KB
constexpr int WG_A_BYTES = 2 * WG_BOX;           // dY: 128 output channels
constexpr int WG_B_BYTES = 4 * WG_BOX;           // X (shifted): 256 input channels
constexpr int WG_STAGE = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 256;
constexpr int WG_THREADS = 256;
constexpr int WG_M = 128, WG_N = 256;

struct WgradParams {
    float* ws;   // [splits][cout][9][cin] fp32 partial sums
    int n_boards, cin, cout, co_tiles, ci_tiles, splits, boards_per_split;
};

// MN-major SWIZZLE_128B shared-memory descriptor: 64-element MN groups `lbo` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, WgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tfull = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int split = t % P.splits;
    t /= P.splits;
    const int tap = t % 9;
    t /= 9;
    const int ci_tile = t % P.ci_tiles, co_tile = t / P.ci_tiles;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int b_lo = split * P.boards_per_split;
    int b_hi = b_lo + P.boards_per_split;
    if (b_hi > P.n_boards) b_hi = P.n_boards;
    const int nk = b_hi > b_lo ? b_hi - b_lo : 0;   // k-blocks: one board (64 pixels) each

    if (warp == 0 && lane == 0) {
        kvu::prefetch_tmap(&tmDY);
        kvu::prefetch_tmap(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; s++) {
            kvu::mbar_init(&full[s], 1);
            kvu::mbar_init(&empty[s], 1);
        }
        kvu::mbar_init(tfull, 1);
        kvu::fence_barrier_init();
    }
    if (warp == 2) kvu::tmem_alloc(tmem_slot, 256);
    kvu::tc_fence_before();
    __syncthreads();
    kvu::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: per board 2 boxes of dY (the tile's 128 output channels) + 4 boxes of shifted X ----------
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nk; kb++) {
            kvu::mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                uint8_t* sa = smem + stage * WG_STAGE;
                kvu::mbar_arrive_expect_tx(&full[stage], WG_STAGE);
                const int b = b_lo + kb;
#pragma unroll
                for (int h = 0; h < 2; h++)
                    kvu::tma_load_4d(sa + h * WG_BOX, &tmDY, &full[stage], co_tile * WG_M + h * 64, 0, 0, b);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    kvu::tma_load_4d(sa + WG_A_BYTES + q * WG_BOX, &tmX, &full[stage], ci_tile * WG_N + q * 64, dx, dy, b);
            }
            __syncwarp();
            if (++stage == WG_STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D[co 128][ci 256] += dY^T (MN-major A) x X (MN-major B), K = 64 pixels per board -----------
        constexpr uint32_t idesc = kvu::make_idesc_bf16(WG_M, WG_N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nk; kb++) {
            kvu::mbar_wait(&full[stage], phase);
            kvu::tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = kvu::smem_u32(smem + stage * WG_STAGE);
                const uint64_t adesc = make_sw128_mnmajor_desc(sa, WG_BOX);
                const uint64_t bdesc = make_sw128_mnmajor_desc(sa + WG_A_BYTES, WG_BOX);
#pragma unroll
                for (int k = 0; k < 4; k++)   // 16 pixels = two 8-row groups = 2048 B per step
                    kvu::umma_bf16(tmem_base, adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (2048 >> 4)), idesc,
                                   (kb | k) != 0);
                kvu::umma_commit(&empty[stage]);
                if (kb == nk - 1) kvu::umma_commit(tfull);
            }
            __syncwarp();
            if (++stage == WG_STAGES) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: TMEM lane = output channel, 256 columns = input channels of this tile ------------------------
        const int q = warp & 3;
        const int co = co_tile * WG_M + q * 32 + lane;
        float* dst = P.ws + (((size_t)split * P.cout + co) * 9 + tap) * P.cin + (size_t)ci_tile * WG_N;
        if (nk > 0) {
            kvu::mbar_wait(tfull, 0);
            kvu::tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < WG_N / 32; c++) {
            uint32_t v[32];
            if (nk > 0) {
                kvu::tmem_ld_32x32(tmem_base + c * 32 + ((uint32_t)(q * 32) << 16), v);
                kvu::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0u;
            }
            float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
            for (int j = 0; j < 8; j++)
                d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                    __uint_as_float(v[4 * j + 3]));
        }
    }
    kvu::tc_fence_before();
    __syncthreads();
    if (warp == 2) kvu::tmem_dealloc(tmem_base, 256);
}

// dW[co][ci][ky][kx] = sum over splits (ascending: deterministic) of ws[s][co][tap][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int cout, int cin, float* __restrict__ dw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t tot = (size_t)cout * 9 * cin;
    if (i >= tot) return;
    const int ci = (int)(i % cin), tap = (int)((i / cin) % 9), co = (int)(i / ((size_t)cin * 9));
    float a = 0.f;
    for (int s = 0; s < splits; s++) a += ws[(size_t)s * tot + i];
    dw[((size_t)co * cin + ci) * 9 + tap] = a;
}

// fp32 [Cout][Cin][3][3] -> bf16 [R][9][K]: plain (R = Cout, K = Cin) or, for dgrad, flipped and transposed
// (R = Cin, K = Cout, tap -> 8 - tap)
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, int cout, int cin, int flip_transpose,
                                    bf16* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t tot = (size_t)cout * 9 * cin;
    if (i >= tot) return;
    if (!flip_transpose) {
        const int ci = (int)(i % cin), tap = (int)((i / cin) % 9), co = (int)(i / ((size_t)cin * 9));
        out[i] = __float2bfloat16(w[((size_t)co * cin + ci) * 9 + tap]);
    } else {
        const int co = (int)(i % cout), tap = (int)((i / cout) % 9), ci = (int)(i / ((size_t)cout * 9));
        out[i] = __float2bfloat16(w[((size_t)co * cin + ci) * 9 + (8 - tap)]);
    }
}

}  // namespace kvt

using namespace kvt;

extern "C" {

int kv_conv3x3_pack(kv_ctx* ctx, const float* d_w, int cout, int cin, int flip_transpose, void* d_out, void* stream) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t tot = (size_t)cout * 9 * cin;
    pack_conv3x3_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_w, cout, cin, flip_transpose,
                                                                                         reinterpret_cast<bf16*>(d_out));
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_conv3x3_fprop(kv_ctx* ctx, const void* d_x, const void* d_w_packed, const float* d_bias, const void* d_residual,
                     void* d_y, int n_boards, int cin, int cout, int relu, void* stream) {
    if (!ctx) return -3;
    if (n_boards <= 0) return 0;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (cout > 512) return kv_fail_msg(ctx, "kv_conv3x3_fprop: cout must be <= 512");
    if (!d_bias) {
        if (!ctx->train_zeros) {
            KV_CUDA(ctx, cudaMalloc(&ctx->train_zeros, 512 * sizeof(float)));
            KV_CUDA(ctx, cudaMemset(ctx->train_zeros, 0, 512 * sizeof(float)));
        }
        d_bias = ctx->train_zeros;
    }
    return kv_conv_launch(ctx, reinterpret_cast<const bf16*>(d_x), reinterpret_cast<const bf16*>(d_w_packed), d_bias,
                          reinterpret_cast<const bf16*>(d_residual), reinterpret_cast<bf16*>(d_y), n_boards, cin, cout, relu,
                          (cudaStream_t)stream);
}

int kv_conv3x3_wgrad(kv_ctx* ctx, const void* d_x, const void* d_dy, float* d_dw, int n_boards, int cin, int cout,
                     void* stream) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_boards <= 0) return kv_fail_msg(ctx, "kv_conv3x3_wgrad: no boards");
    if (cin % WG_N || cout % WG_M) return kv_fail_msg(ctx, "kv_conv3x3_wgrad: cin must be a multiple of 256, cout of 128");
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->wgrad_attr_done) {
        KV_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        ctx->wgrad_attr_done = true;
    }
    WgradParams P;
    P.n_boards = n_boards;
    P.cin = cin;
    P.cout = cout;
    P.co_tiles = cout / WG_M;
    P.ci_tiles = cin / WG_N;
    const int tiles = P.co_tiles * P.ci_tiles * 9;
    int splits = ctx->sm_count / tiles;
    if (splits < 1) splits = 1;
    if (splits > n_boards) splits = n_boards;
    P.splits = splits;
    P.boards_per_split = (n_boards + splits - 1) / splits;
    const size_t need = (size_t)splits * cout * 9 * cin;
    if (ctx->train_ws_floats < need) {
        KV_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->train_ws) cudaFree(ctx->train_ws);
        ctx->train_ws = nullptr;
        ctx->train_ws_floats = 0;
        KV_CUDA(ctx, cudaMalloc(&ctx->train_ws, need * sizeof(float)));
        ctx->train_ws_floats = need;
    }
    P.ws = ctx->train_ws;
    CUtensorMap mdy, mx;
    if (int rc = kv_make_act_map(ctx, &mdy, const_cast<void*>(d_dy), cout, n_boards, 1)) return rc;
    if (int rc = kv_make_act_map(ctx, &mx, const_cast<void*>(d_x), cin, n_boards, 1)) return rc;
    {
        KvTimed t_(ctx, KVK_TRAIN_WGRAD, st);
        conv3x3_wgrad_kernel<<<tiles * splits, WG_THREADS, WG_SMEM, st>>>(mdy, mx, P);
        KV_LAUNCH_CHECK(ctx);
        const size_t tot = (size_t)cout * 9 * cin;
        wgrad_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ctx->train_ws, splits, cout, cin, d_dw);
        KV_LAUNCH_CHECK(ctx);
    }
    return 0;
}

}  // extern "C"
