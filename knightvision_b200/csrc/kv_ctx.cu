// kv_ctx.cu — context lifetime and the host-buffer entry points of libkv_b200.so
#include <cstring>

#include "kv_internal.h"

std::string g_kv_create_error;

void kv_net_destroy(kv_ctx* ctx);
void kv_mcts_destroy(kv_ctx* ctx);

int kv_stage_reserve(kv_ctx* ctx, size_t host_bytes, size_t dev_bytes) {
    if (ctx->h_stage_bytes < host_bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
        KV_CUDA(ctx, cudaMallocHost(&ctx->h_stage, host_bytes));
        ctx->h_stage_bytes = host_bytes;
    }
    if (ctx->d_stage_bytes < dev_bytes) {
        if (ctx->d_stage) cudaFree(ctx->d_stage);
        ctx->d_stage = nullptr;
        ctx->d_stage_bytes = 0;
        KV_CUDA(ctx, cudaMalloc(&ctx->d_stage, dev_bytes));
        ctx->d_stage_bytes = dev_bytes;
    }
    return 0;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int kv_abi_version(void) { return 1; }

int kv_create(int device, kv_ctx** out) {
    if (!out) return -3;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_kv_create_error = std::string("kv_create: no CUDA device (") + cudaGetErrorString(e) +
                            "); libkv_b200 has no CPU fallback";
        return -1;
    }
    if (device < 0 || device >= count) {
        g_kv_create_error = "kv_create: bad device index";
        return -2;
    }
    KV_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    KV_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        g_kv_create_error = "kv_create: this library is built for sm_100a (B200) only";
        return -4;
    }
    kv_ctx* ctx = new kv_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return 0;
}

void kv_destroy(kv_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    kv_net_destroy(ctx);
    kv_mcts_destroy(ctx);
    for (int l = 0; l < 8; l++)
        if (ctx->perft_buf[l]) cudaFree(ctx->perft_buf[l]);
    if (ctx->perft_counter) cudaFree(ctx->perft_counter);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    if (ctx->train_ws) cudaFree(ctx->train_ws);
    if (ctx->train_zeros) cudaFree(ctx->train_zeros);
    if (ctx->bn_ws) cudaFree(ctx->bn_ws);
    for (auto& p : ctx->ev_live) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto& p : ctx->ev_free) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    delete ctx;
}

const char* kv_last_error(kv_ctx* ctx) { return ctx ? ctx->err.c_str() : g_kv_create_error.c_str(); }
int kv_sm_count(kv_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t kv_launch_count(kv_ctx* ctx) { return ctx ? ctx->launches : 0; }

int kv_profile_enable(kv_ctx* ctx, int on) {
    if (!ctx) return -3;
    ctx->profiling = on != 0;
    return 0;
}

// Sums the CUDA-event durations recorded since the last call into ms[KVK_COUNT] / n[KVK_COUNT] (synchronises).
int kv_profile_read(kv_ctx* ctx, double* ms, uint64_t* n, int cap) {
    if (!ctx) return -3;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    KV_CUDA(ctx, cudaDeviceSynchronize());
    for (auto& p : ctx->ev_live) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) {
            ctx->prof_ms[p.kind] += t;
            ctx->prof_n[p.kind] += 1;
        }
        ctx->ev_free.push_back(p);
    }
    ctx->ev_live.clear();
    for (int k = 0; k < KVK_COUNT && k < cap; k++) {
        if (ms) ms[k] = ctx->prof_ms[k];
        if (n) n[k] = ctx->prof_n[k];
        ctx->prof_ms[k] = 0;
        ctx->prof_n[k] = 0;
    }
    return KVK_COUNT;
}

int kv_movegen_host(kv_ctx* ctx, uint64_t* h_lines, int n, uint16_t* h_moves, int stride, int32_t* h_counts,
                    int32_t* h_flags) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t lb = align256((size_t)n * 128), mb = align256((size_t)n * stride * 2), cb = align256((size_t)n * 4);
    const size_t total = lb + mb + 2 * cb;
    if (int rc = kv_stage_reserve(ctx, total, total)) return rc;
    char* h = (char*)ctx->h_stage;
    char* d = (char*)ctx->d_stage;
    memcpy(h, h_lines, (size_t)n * 128);
    KV_CUDA(ctx, cudaMemcpyAsync(d, h, (size_t)n * 128, cudaMemcpyHostToDevice, 0));
    if (int rc = kv_movegen(ctx, (uint64_t*)d, n, (uint16_t*)(d + lb), stride, (int32_t*)(d + lb + mb),
                            (int32_t*)(d + lb + mb + cb), nullptr))
        return rc;
    KV_CUDA(ctx, cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, 0));
    KV_CUDA(ctx, cudaStreamSynchronize(0));
    memcpy(h_lines, h, (size_t)n * 128);
    memcpy(h_counts, h + lb + mb, (size_t)n * 4);
    memcpy(h_flags, h + lb + mb + cb, (size_t)n * 4);
    // only the first count entries of each list are defined
    const uint16_t* src = (const uint16_t*)(h + lb);
    for (int i = 0; i < n; i++) {
        int c = h_counts[i] < stride ? h_counts[i] : stride;
        memcpy(h_moves + (size_t)i * stride, src + (size_t)i * stride, (size_t)c * 2);
        memset(h_moves + (size_t)i * stride + c, 0, (size_t)(stride - c) * 2);
    }
    return 0;
}

int kv_make_moves_host(kv_ctx* ctx, uint64_t* h_lines, int n, const uint16_t* h_moves) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t lb = align256((size_t)n * 128), mb = align256((size_t)n * 2);
    if (int rc = kv_stage_reserve(ctx, lb + mb, lb + mb)) return rc;
    char* h = (char*)ctx->h_stage;
    char* d = (char*)ctx->d_stage;
    memcpy(h, h_lines, (size_t)n * 128);
    memcpy(h + lb, h_moves, (size_t)n * 2);
    KV_CUDA(ctx, cudaMemcpyAsync(d, h, lb + mb, cudaMemcpyHostToDevice, 0));
    if (int rc = kv_make_moves(ctx, (uint64_t*)d, n, (const uint16_t*)(d + lb), nullptr)) return rc;
    KV_CUDA(ctx, cudaMemcpyAsync(h, d, (size_t)n * 128, cudaMemcpyDeviceToHost, 0));
    KV_CUDA(ctx, cudaStreamSynchronize(0));
    memcpy(h_lines, h, (size_t)n * 128);
    return 0;
}

int kv_attacked_host(kv_ctx* ctx, const uint64_t* h_lines, int n, uint64_t* h_masks) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t lb = align256((size_t)n * 128), mb = align256((size_t)n * 8);
    if (int rc = kv_stage_reserve(ctx, lb + mb, lb + mb)) return rc;
    char* h = (char*)ctx->h_stage;
    char* d = (char*)ctx->d_stage;
    memcpy(h, h_lines, (size_t)n * 128);
    KV_CUDA(ctx, cudaMemcpyAsync(d, h, (size_t)n * 128, cudaMemcpyHostToDevice, 0));
    if (int rc = kv_attacked(ctx, (const uint64_t*)d, n, (uint64_t*)(d + lb), nullptr)) return rc;
    KV_CUDA(ctx, cudaMemcpyAsync(h + lb, d + lb, (size_t)n * 8, cudaMemcpyDeviceToHost, 0));
    KV_CUDA(ctx, cudaStreamSynchronize(0));
    memcpy(h_masks, h + lb, (size_t)n * 8);
    return 0;
}

int kv_perft_host(kv_ctx* ctx, const uint64_t* h_roots, int n, int depth, uint64_t* h_out, int chunk) {
    if (!ctx) return -3;
    if (n <= 0) return 0;
    KV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t lb = align256((size_t)n * 128), ob = align256((size_t)n * 64);
    if (int rc = kv_stage_reserve(ctx, lb + ob, lb + ob)) return rc;
    char* h = (char*)ctx->h_stage;
    char* d = (char*)ctx->d_stage;
    memcpy(h, h_roots, (size_t)n * 128);
    KV_CUDA(ctx, cudaMemcpyAsync(d, h, (size_t)n * 128, cudaMemcpyHostToDevice, 0));
    if (int rc = kv_perft(ctx, (const uint64_t*)d, n, depth, (uint64_t*)(d + lb), chunk, nullptr)) return rc;
    KV_CUDA(ctx, cudaMemcpyAsync(h + lb, d + lb, (size_t)n * 64, cudaMemcpyDeviceToHost, 0));
    KV_CUDA(ctx, cudaStreamSynchronize(0));
    memcpy(h_out, h + lb, (size_t)n * 64);
    return 0;
}

}  // extern "C"
