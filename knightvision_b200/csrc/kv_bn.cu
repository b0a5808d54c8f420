// kv_bn.cu — train-mode BatchNorm + ReLU (+ residual add) of the policy/value tower on NHWC bf16 activations.
//
// Reference: ai/model.py:19-25,58-59 (`F.relu(bn(conv(x)))`, `F.relu(bn2(conv2(t)) + x)`) as scripts/train.py:158-181
// runs it in training mode (batch statistics, running-stat update with momentum 0.1, eps 1e-5).  These are
// HBM-bound passes over [rows = boards * 64][C] bf16 tensors, so the design rule is bytes, not flops:
//   forward   stats pass (read z) -> finalize (mean, rstd, running stats) -> apply pass (read z [+ residual], write y)
//   backward  reduce pass (read dy, y, z -> sum g, sum g*zhat with g = dy * [y > 0]) -> finalize (dgamma, dbeta)
//             -> apply pass (read dy, y, z; write dz [and g for the skip connection])
// A thread owns 8 consecutive channels (one 16 B load/store) of a row; the C/8 threads of a row group read a whole
// row contiguously, a 256-thread CTA covers 2 048 / C rows per iteration.  Reductions are two-stage and deterministic:
// per-CTA partial sums in fp32 (fixed intra-CTA order), then one thread per channel adds the partials in CTA order in
// fp64.  Grid = 4 CTAs per SM.
#include <cuda_bf16.h>

#include <cstdint>

#include "kv_internal.h"

using bf16 = __nv_bfloat16;

namespace kvb {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_C = 1024;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        f[2 * i] = __low2float(h[i]);
        f[2 * i + 1] = __high2float(h[i]);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}

__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// MODE 0: (z, z^2)   MODE 1: (g, g * zhat) with g = dy * [y > 0] (relu) or dy   MODE 2: (x, -) plain column sums
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_reduce_kernel(const bf16* __restrict__ a, const bf16* __restrict__ yv,
                                                               const bf16* __restrict__ zv, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, int rows, int C, int relu,
                                                               float* __restrict__ partial /*[grid][2][C]*/) {
    __shared__ float red[2][BN_THREADS][9];   // +1 padding against bank conflicts
    const int tpr = C >> 3;                   // threads per row
    const int rpi = BN_THREADS / tpr;         // rows per CTA iteration
    const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
    const int c0 = cg * 8;
    float s0[8], s1[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s0[i] = 0.f;
        s1[i] = 0.f;
        mu[i] = MODE == 1 ? mean[c0 + i] : 0.f;
        rs[i] = MODE == 1 ? rstd[c0 + i] : 0.f;
    }
    // contiguous slab of rows per CTA (deterministic partition)
    const int per = (rows + gridDim.x - 1) / gridDim.x;
    const int r_lo = blockIdx.x * per;
    const int r_hi = r_lo + per < rows ? r_lo + per : rows;
    // four rows per trip: all loads of a trip are issued before the first use (bytes in flight hide the HBM latency)
    constexpr int U = 4;
    for (int r = r_lo + rl; r < r_hi; r += U * rpi) {
        uint4 va[U], vz[U], vy[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int rr = r + u * rpi;
            if (rr < r_hi) {
                const size_t off = (size_t)rr * C + c0;
                va[u] = __ldg(reinterpret_cast<const uint4*>(a + off));
                if (MODE == 1) {
                    vz[u] = __ldg(reinterpret_cast<const uint4*>(zv + off));
                    if (relu) vy[u] = __ldg(reinterpret_cast<const uint4*>(yv + off));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (r + u * rpi >= r_hi) break;
            float x[8];
            unpack8(va[u], x);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    s0[i] += x[i];
                    s1[i] += x[i] * x[i];
                }
            } else if (MODE == 1) {
                float y[8], z[8];
                unpack8(vz[u], z);
                if (relu) unpack8(vy[u], y);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float g = (!relu || y[i] > 0.f) ? x[i] : 0.f;
                    s0[i] += g;
                    s1[i] += g * ((z[i] - mu[i]) * rs[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) s0[i] += x[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        red[0][threadIdx.x][i] = s0[i];
        red[1][threadIdx.x][i] = s1[i];
    }
    __syncthreads();
    // one thread per (which, channel): add the CTA's row lanes in fixed order
    for (int j = threadIdx.x; j < 2 * C; j += BN_THREADS) {
        const int which = j / C, c = j % C;
        float t = 0.f;
        for (int k = 0; k < rpi; k++) t += red[which][k * tpr + (c >> 3)][c & 7];
        partial[((size_t)blockIdx.x * 2 + which) * C + c] = t;
    }
}

// Second reduction stage.  CTA = 8 channels x 32 slices of the partial list; slice p adds partials p, p + 32, ... in
// fp64 (all its loads issued first), then the 32 slice sums are added in slice order: deterministic, short dependency
// chains, and C / 8 CTAs instead of a handful.
constexpr int FIN_THREADS = 256;
constexpr int FIN_CH = 8, FIN_SLICES = FIN_THREADS / FIN_CH;   // 32
constexpr int FIN_MAX_TRIPS = 24;                              // up to 32 * 24 = 768 partials (grid <= 4 * 148 = 592)
__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int nblk, int C, int c, double& s, double& q,
                                             double (*sh)[2][FIN_CH]) {
    const int part = threadIdx.x / FIN_CH, cl = threadIdx.x % FIN_CH;
    float va[FIN_MAX_TRIPS], vb[FIN_MAX_TRIPS];
#pragma unroll
    for (int u = 0; u < FIN_MAX_TRIPS; u++) {
        const int k = part + FIN_SLICES * u;
        const bool ok = c < C && k < nblk;
        va[u] = ok ? __ldg(partial + ((size_t)k * 2 + 0) * C + c) : 0.f;
        vb[u] = ok ? __ldg(partial + ((size_t)k * 2 + 1) * C + c) : 0.f;
    }
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int u = 0; u < FIN_MAX_TRIPS; u++) {
        a += (double)va[u];
        b += (double)vb[u];
    }
    sh[part][0][cl] = a;
    sh[part][1][cl] = b;
    __syncthreads();
    s = 0.0;
    q = 0.0;
    for (int p = 0; p < FIN_SLICES; p++) {
        s += sh[p][0][cl];
        q += sh[p][1][cl];
    }
}

// forward finalize: mean / rstd (biased variance) + running-stat update (unbiased variance), as torch.nn.BatchNorm2d
__global__ void __launch_bounds__(FIN_THREADS) bn_fwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C,
                                                                      int rows, float momentum, float eps,
                                                                      float* __restrict__ save_mean,
                                                                      float* __restrict__ save_rstd,
                                                                      float* __restrict__ running_mean,
                                                                      float* __restrict__ running_var) {
    __shared__ double sh[FIN_SLICES][2][FIN_CH];
    const int c = blockIdx.x * FIN_CH + (threadIdx.x % FIN_CH);
    double s, q;
    sum_partials(partial, nblk, C, c, s, q, sh);
    if (threadIdx.x >= FIN_CH || c >= C) return;
    const double n = (double)rows;
    const double m = s / n;
    double var = q / n - m * m;
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)m;
    save_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    if (running_var) {
        const double unb = rows > 1 ? var * n / (n - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// sums of the partials -> out0 / out1 (either may be null)
__global__ void __launch_bounds__(FIN_THREADS) bn_sum_partials_kernel(const float* __restrict__ partial, int nblk, int C,
                                                                      float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ double sh[FIN_SLICES][2][FIN_CH];
    const int c = blockIdx.x * FIN_CH + (threadIdx.x % FIN_CH);
    double s, q;
    sum_partials(partial, nblk, C, c, s, q, sh);
    if (threadIdx.x >= FIN_CH || c >= C) return;
    if (out0) out0[c] = (float)s;
    if (out1) out1[c] = (float)q;
}

// The apply kernels run with a grid whose thread count is a multiple of C / 8, so a thread keeps its 8 channels for
// the whole grid-stride loop: the per-channel parameters are loaded once, and two vectors are in flight per trip.
// y = [relu](gamma * (z - mean) * rstd + beta [+ residual])
__global__ void __launch_bounds__(BN_THREADS) bn_apply_fwd_kernel(const bf16* __restrict__ z, const bf16* __restrict__ res,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, size_t n8, int C, int relu,
                                                                  bf16* __restrict__ y) {
    const size_t stride = (size_t)gridDim.x * BN_THREADS;
    const size_t i0 = (size_t)blockIdx.x * BN_THREADS + threadIdx.x;
    const int c0 = (int)(i0 % (size_t)(C >> 3)) * 8;
    float sc[8], sh[8];
    {
        float ga[8], be[8], mu[8], rs[8];
        ld8(gamma + c0, ga);
        ld8(beta + c0, be);
        ld8(mean + c0, mu);
        ld8(rstd + c0, rs);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            sc[k] = rs[k] * ga[k];
            sh[k] = be[k] - mu[k] * sc[k];
        }
    }
    for (size_t i = i0; i < n8; i += 2 * stride) {
        const size_t j = i + stride;
        const bool two = j < n8;
        uint4 vx[2], vr[2];
        vx[0] = __ldg(reinterpret_cast<const uint4*>(z) + i);
        if (two) vx[1] = __ldg(reinterpret_cast<const uint4*>(z) + j);
        if (res) {
            vr[0] = __ldg(reinterpret_cast<const uint4*>(res) + i);
            if (two) vr[1] = __ldg(reinterpret_cast<const uint4*>(res) + j);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (u == 1 && !two) break;
            float x[8], r[8], o[8];
            unpack8(vx[u], x);
            if (res) unpack8(vr[u], r);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                float v = x[k] * sc[k] + sh[k];
                if (res) v += r[k];
                o[k] = relu ? fmaxf(v, 0.f) : v;
            }
            reinterpret_cast<uint4*>(y)[u ? j : i] = pack8(o);
        }
    }
}

// dz = gamma * rstd * (g - dbeta / N - zhat * dgamma / N), g = dy * [y > 0]; dres = g (gradient of the skip input)
__global__ void __launch_bounds__(BN_THREADS) bn_apply_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ yv,
                                                                  const bf16* __restrict__ zv,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  const float* __restrict__ dgamma,
                                                                  const float* __restrict__ dbeta, size_t n8, int C, int relu,
                                                                  float inv_n, bf16* __restrict__ dz, bf16* __restrict__ dres) {
    const size_t stride = (size_t)gridDim.x * BN_THREADS;
    const size_t i0 = (size_t)blockIdx.x * BN_THREADS + threadIdx.x;
    const int c0 = (int)(i0 % (size_t)(C >> 3)) * 8;
    float ga[8], mu[8], rs[8], dg[8], db[8];
    ld8(gamma + c0, ga);
    ld8(mean + c0, mu);
    ld8(rstd + c0, rs);
    ld8(dgamma + c0, dg);
    ld8(dbeta + c0, db);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        ga[k] = ga[k] * rs[k];      // gamma * rstd
        dg[k] = dg[k] * inv_n;      // dgamma / N
        db[k] = db[k] * inv_n;      // dbeta / N
    }
    for (size_t i = i0; i < n8; i += 2 * stride) {
        const size_t j = i + stride;
        const bool two = j < n8;
        uint4 vd[2], vz[2], vy[2];
        vd[0] = __ldg(reinterpret_cast<const uint4*>(dy) + i);
        vz[0] = __ldg(reinterpret_cast<const uint4*>(zv) + i);
        if (relu) vy[0] = __ldg(reinterpret_cast<const uint4*>(yv) + i);
        if (two) {
            vd[1] = __ldg(reinterpret_cast<const uint4*>(dy) + j);
            vz[1] = __ldg(reinterpret_cast<const uint4*>(zv) + j);
            if (relu) vy[1] = __ldg(reinterpret_cast<const uint4*>(yv) + j);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (u == 1 && !two) break;
            float d[8], y[8], z[8], o[8], g[8];
            unpack8(vd[u], d);
            unpack8(vz[u], z);
            if (relu) unpack8(vy[u], y);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                g[k] = (!relu || y[k] > 0.f) ? d[k] : 0.f;
                const float zh = (z[k] - mu[k]) * rs[k];
                o[k] = ga[k] * (g[k] - db[k] - zh * dg[k]);
            }
            reinterpret_cast<uint4*>(dz)[u ? j : i] = pack8(o);
            if (dres) reinterpret_cast<uint4*>(dres)[u ? j : i] = pack8(g);
        }
    }
}

}  // namespace kvb

using namespace kvb;

static int bn_grid(kv_ctx* ctx, int rows, int C) {
    const int rpi = BN_THREADS / (C >> 3);
    int g = ctx->sm_count * 4;
    if (g > FIN_SLICES * FIN_MAX_TRIPS) g = FIN_SLICES * FIN_MAX_TRIPS;   // what the second stage reads
    const int need = (rows + rpi - 1) / rpi;
    if (g > need) g = need;
    return g < 1 ? 1 : g;
}
static int bn_check(kv_ctx* ctx, int rows, int C) {
    if (rows < 1) return kv_fail_msg(ctx, "kv_bn: no rows");
    if (C % 8 || C > BN_MAX_C || BN_THREADS % (C >> 3)) return kv_fail_msg(ctx, "kv_bn: C must be 64, 128, 256, 512 or 1024");
    return 0;
}
static int bn_ws(kv_ctx* ctx, int grid, int C) {
    const size_t need = (size_t)grid * 2 * C;
    if (ctx->bn_ws_floats < need) {
        KV_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->bn_ws) cudaFree(ctx->bn_ws);
        ctx->bn_ws = nullptr;
        ctx->bn_ws_floats = 0;
        KV_CUDA(ctx, cudaMalloc(&ctx->bn_ws, need * sizeof(float)));
        ctx->bn_ws_floats = need;
    }
    return 0;
}

extern "C" {

int kv_bn_relu_fwd(kv_ctx* ctx, const void* d_z, const void* d_residual, const float* d_gamma, const float* d_beta,
                   float* d_running_mean, float* d_running_var, float momentum, float eps, void* d_y, float* d_save_mean,
                   float* d_save_rstd, int rows, int C, int relu, void* stream) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = bn_check(ctx, rows, C)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = bn_grid(ctx, rows, C);
    if (int rc = bn_ws(ctx, grid, C)) return rc;
    KvTimed t_(ctx, KVK_TRAIN_BN, st);
    bn_reduce_kernel<0><<<grid, BN_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(d_z), nullptr, nullptr, nullptr, nullptr,
                                                     rows, C, 0, ctx->bn_ws);
    KV_LAUNCH_CHECK(ctx);
    bn_fwd_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, FIN_THREADS, 0, st>>>(ctx->bn_ws, grid, C, rows, momentum, eps, d_save_mean, d_save_rstd,
                                                            d_running_mean, d_running_var);
    KV_LAUNCH_CHECK(ctx);
    const size_t n8 = (size_t)rows * C / 8;
    size_t ag = (n8 + BN_THREADS - 1) / BN_THREADS;
    if (ag > (size_t)ctx->sm_count * 8) ag = (size_t)ctx->sm_count * 8;
    bn_apply_fwd_kernel<<<(unsigned)ag, BN_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(d_z),
                                                             reinterpret_cast<const bf16*>(d_residual), d_gamma, d_beta,
                                                             d_save_mean, d_save_rstd, n8, C, relu, reinterpret_cast<bf16*>(d_y));
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_bn_relu_bwd(kv_ctx* ctx, const void* d_dy, const void* d_y, const void* d_z, const float* d_gamma,
                   const float* d_save_mean, const float* d_save_rstd, void* d_dz, void* d_dres, float* d_dgamma,
                   float* d_dbeta, int rows, int C, int relu, void* stream) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = bn_check(ctx, rows, C)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = bn_grid(ctx, rows, C);
    if (int rc = bn_ws(ctx, grid, C)) return rc;
    KvTimed t_(ctx, KVK_TRAIN_BN, st);
    bn_reduce_kernel<1><<<grid, BN_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(d_dy), reinterpret_cast<const bf16*>(d_y),
                                                     reinterpret_cast<const bf16*>(d_z), d_save_mean, d_save_rstd, rows, C, relu,
                                                     ctx->bn_ws);
    KV_LAUNCH_CHECK(ctx);
    bn_sum_partials_kernel<<<(C + FIN_CH - 1) / FIN_CH, FIN_THREADS, 0, st>>>(ctx->bn_ws, grid, C, d_dbeta, d_dgamma);
    KV_LAUNCH_CHECK(ctx);
    const size_t n8 = (size_t)rows * C / 8;
    size_t ag = (n8 + BN_THREADS - 1) / BN_THREADS;
    if (ag > (size_t)ctx->sm_count * 8) ag = (size_t)ctx->sm_count * 8;
    bn_apply_bwd_kernel<<<(unsigned)ag, BN_THREADS, 0, st>>>(
        reinterpret_cast<const bf16*>(d_dy), reinterpret_cast<const bf16*>(d_y), reinterpret_cast<const bf16*>(d_z), d_gamma,
        d_save_mean, d_save_rstd, d_dgamma, d_dbeta, n8, C, relu, 1.0f / (float)rows, reinterpret_cast<bf16*>(d_dz),
        reinterpret_cast<bf16*>(d_dres));
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

int kv_channel_sum(kv_ctx* ctx, const void* d_x, float* d_out, int rows, int C, void* stream) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = bn_check(ctx, rows, C)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = bn_grid(ctx, rows, C);
    if (int rc = bn_ws(ctx, grid, C)) return rc;
    KvTimed t_(ctx, KVK_TRAIN_BN, st);
    bn_reduce_kernel<2><<<grid, BN_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(d_x), nullptr, nullptr, nullptr, nullptr, rows,
                                                     C, 0, ctx->bn_ws);
    KV_LAUNCH_CHECK(ctx);
    bn_sum_partials_kernel<<<(C + FIN_CH - 1) / FIN_CH, FIN_THREADS, 0, st>>>(ctx->bn_ws, grid, C, d_out, nullptr);
    KV_LAUNCH_CHECK(ctx);
    return 0;
}

}  // extern "C"
