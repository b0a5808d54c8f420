// kv_heads.cuh — policy / value head device functions (ai/model.py:64-73), shared by the full-logits forward
// (kv_net.cu) and the search-mode evaluator that only needs the legal moves' logits (kv_mcts.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace kvn {
using bf16 = __nv_bfloat16;

// wh [3][C] fp32 (policy ch0, policy ch1, value; BN folded), bh [3].  One CTA (>= 256 threads use the first 256)
// per board.  feat out: hp[128] (index c*64 + pixel, torch.flatten order of [2,8,8]) and hv[64], both after relu.
// swh: shared-memory staging for wh (3*C floats).  Four consecutive lanes share a pixel and read four consecutive
// 16 B chunks of its row (64 B contiguous per pixel per request: full sectors).
__device__ __forceinline__ void head_features(const bf16* __restrict__ act, int C, const float* __restrict__ wh,
                                              const float* __restrict__ bh, float* hp, float* hv, float* swh) {
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) swh[i] = __ldg(wh + i);
    __syncthreads();
    if (threadIdx.x < 256) {
        const int px = threadIdx.x >> 2, part = threadIdx.x & 3;
        const bf16* row = act + (size_t)px * C;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int c = part * 8; c < C; c += 32) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + c));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float4 w0a = *reinterpret_cast<const float4*>(swh + c), w0b = *reinterpret_cast<const float4*>(swh + c + 4);
            const float4 w1a = *reinterpret_cast<const float4*>(swh + C + c), w1b = *reinterpret_cast<const float4*>(swh + C + c + 4);
            const float4 w2a = *reinterpret_cast<const float4*>(swh + 2 * C + c), w2b = *reinterpret_cast<const float4*>(swh + 2 * C + c + 4);
            const float a0 = __low2float(h[0]), a1 = __high2float(h[0]), a2 = __low2float(h[1]), a3 = __high2float(h[1]);
            const float a4 = __low2float(h[2]), a5 = __high2float(h[2]), a6 = __low2float(h[3]), a7 = __high2float(h[3]);
            s0 += a0 * w0a.x + a1 * w0a.y + a2 * w0a.z + a3 * w0a.w + a4 * w0b.x + a5 * w0b.y + a6 * w0b.z + a7 * w0b.w;
            s1 += a0 * w1a.x + a1 * w1a.y + a2 * w1a.z + a3 * w1a.w + a4 * w1b.x + a5 * w1b.y + a6 * w1b.z + a7 * w1b.w;
            s2 += a0 * w2a.x + a1 * w2a.y + a2 * w2a.z + a3 * w2a.w + a4 * w2b.x + a5 * w2b.y + a6 * w2b.z + a7 * w2b.w;
        }
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, m);
            s1 += __shfl_xor_sync(0xffffffffu, s1, m);
            s2 += __shfl_xor_sync(0xffffffffu, s2, m);
        }
        if (part == 0) {
            hp[px] = fmaxf(s0 + bh[0], 0.f);
            hp[64 + px] = fmaxf(s1 + bh[1], 0.f);
            hv[px] = fmaxf(s2 + bh[2], 0.f);
        }
    }
}

__device__ __forceinline__ float value_mlp(const float* hv, const float* __restrict__ w1, const float* __restrict__ b1,
                                           const float* __restrict__ w2, const float* __restrict__ b2, float* red) {
    // value_fc1 64->512 + relu, value_fc2 512->1, tanh (ai/model.py:70-73).  w1 is TRANSPOSED here: w1[i * 512 + j] is
    // value_fc1.weight[j][i], so the 32 lanes of a warp read 128 contiguous bytes per input i.  The sum keeps the
    // association of the row-major version (four products added left to right, then added to the accumulator).
    float part = 0.f;
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
        float a = __ldg(b1 + j);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float wx = __ldg(w1 + (4 * i + 0) * 512 + j), wy = __ldg(w1 + (4 * i + 1) * 512 + j);
            const float wz = __ldg(w1 + (4 * i + 2) * 512 + j), ww = __ldg(w1 + (4 * i + 3) * 512 + j);
            a += wx * hv[4 * i] + wy * hv[4 * i + 1] + wz * hv[4 * i + 2] + ww * hv[4 * i + 3];
        }
        part += fmaxf(a, 0.f) * __ldg(w2 + j);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += red[w];
    return tanhf(tot + __ldg(b2));
}


}  // namespace kvn
