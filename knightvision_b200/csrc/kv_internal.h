// kv_internal.h — context and launch helpers shared by the .cu files of libkv_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/kv_b200.h"

struct kv_net;    // kv_net.cu
struct kv_mcts;   // kv_mcts.cu

// kernel categories for the optional per-kernel CUDA-event timing (kv_profile_*)
enum KvKernel : int {
    KVK_MOVEGEN = 0, KVK_MAKE_MOVES, KVK_PERFT_EXPAND, KVK_PERFT_LEAF, KVK_ENCODE,
    KVK_NET_STEM, KVK_NET_CONV, KVK_NET_HEAD, KVK_MCTS_SELECT, KVK_MCTS_EXPAND, KVK_MCTS_MISC, KVK_TRAIN_WGRAD, KVK_TRAIN_BN, KVK_COUNT
};

struct KvEventPair {
    cudaEvent_t a, b;
    int kind;
};

struct kv_ctx {
    bool profiling = false;
    std::vector<KvEventPair> ev_live;
    std::vector<KvEventPair> ev_free;
    double prof_ms[KVK_COUNT] = {0};
    uint64_t prof_n[KVK_COUNT] = {0};
    int device = 0;
    int sm_count = 148;
    std::string err;
    uint64_t launches = 0;
    // perft level buffers (lazily allocated)
    uint64_t* perft_buf[8] = {nullptr};
    size_t perft_cap = 0;   // boards per level buffer (levels 1..7)
    size_t perft_roots_cap = 0;   // boards in perft_buf[0]
    uint32_t* perft_counter = nullptr;
    // pinned + device staging for the *_host entry points
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    void* d_stage = nullptr;
    size_t d_stage_bytes = 0;
    kv_net* net = nullptr;
    kv_mcts* mcts = nullptr;
    // training operators (kv_train.cu): split-K workspace of the weight-gradient kernel, zero bias
    float* train_ws = nullptr;
    size_t train_ws_floats = 0;
    float* train_zeros = nullptr;
    float* bn_ws = nullptr;          // per-CTA partial sums of the BatchNorm reductions (kv_bn.cu)
    size_t bn_ws_floats = 0;
    bool conv_attr_done = false;     // cudaFuncSetAttribute(max dynamic smem) is per device: once per context
    bool wgrad_attr_done = false;
};

extern std::string g_kv_create_error;

inline int kv_fail(kv_ctx* ctx, const char* what, cudaError_t e, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (ctx) ctx->err = buf;
    else g_kv_create_error = buf;
    return -1;
}
inline int kv_fail_msg(kv_ctx* ctx, const char* msg) {
    if (ctx) ctx->err = msg;
    else g_kv_create_error = msg;
    return -2;
}

#define KV_CUDA(ctx, call)                                                  \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return kv_fail(ctx, #call, e__, __FILE__, __LINE__); \
    } while (0)

#define KV_LAUNCH_CHECK(ctx)                                                \
    do {                                                                    \
        (ctx)->launches++;                                                  \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return kv_fail(ctx, "kernel launch", e__, __FILE__, __LINE__); \
    } while (0)

int kv_stage_reserve(kv_ctx* ctx, size_t host_bytes, size_t dev_bytes);

// RAII bracket: records an event pair around a launch when profiling is on (no-op otherwise)
struct KvTimed {
    kv_ctx* ctx;
    cudaStream_t st;
    KvEventPair p;
    bool on;
    KvTimed(kv_ctx* c, int kind, cudaStream_t s) : ctx(c), st(s), on(c->profiling) {
        if (!on) return;
        if (!ctx->ev_free.empty()) {
            p = ctx->ev_free.back();
            ctx->ev_free.pop_back();
        } else {
            cudaEventCreate(&p.a);
            cudaEventCreate(&p.b);
        }
        p.kind = kind;
        cudaEventRecord(p.a, st);
    }
    ~KvTimed() {
        if (!on) return;
        cudaEventRecord(p.b, st);
        ctx->ev_live.push_back(p);
    }
};
