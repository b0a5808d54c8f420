// kv_warp.cuh — warp-collective wrappers (full warp and W-lane groups) used by the rules and tree-search kernels.
//
// On the device these are the sm_100a intrinsics.  When KV_HOST_EMU is defined (tests/simt_emu only —
// a CI harness that runs the *kernel source* lane-by-lane on the CPU so the integer kernels can be
// checked against the oracle on a box without a GPU) they are routed to a 32-fiber lock-step emulator.
// The product library is never built with KV_HOST_EMU.
#pragma once
#include <cstdint>

#ifdef KV_HOST_EMU
#include <cstring>
#define KV_DEV inline
#define KV_DEVFN
namespace kvemu {
uint64_t collective_shfl(uint64_t v, int src);
uint32_t collective_ballot(bool p);
void collective_sync();
int lane_id();
}  // namespace kvemu
namespace kv {
KV_DEV int popc32(uint32_t x) { return __builtin_popcount(x); }
KV_DEV int popc64(uint64_t x) { return __builtin_popcountll(x); }
KV_DEV int ctz64(uint64_t x) { return __builtin_ctzll(x); }        // x != 0
KV_DEV int msb64(uint64_t x) { return 63 - __builtin_clzll(x); }   // x != 0
KV_DEV int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
KV_DEV uint64_t shfl64(uint64_t v, int src) { return kvemu::collective_shfl(v, src); }
KV_DEV int shfl32(int v, int src) { return (int)(int64_t)kvemu::collective_shfl((uint64_t)(int64_t)v, src); }
KV_DEV float shflf(float v, int src) {
    uint32_t u; memcpy(&u, &v, 4);
    u = (uint32_t)kvemu::collective_shfl(u, src);
    float r; memcpy(&r, &u, 4); return r;
}
KV_DEV uint32_t ballot(bool p) { return kvemu::collective_ballot(p); }
KV_DEV void syncwarp() { kvemu::collective_sync(); }
KV_DEV int shfl_up32(int v, int delta, int lane) {
    int r = (int)(int64_t)kvemu::collective_shfl((uint64_t)(int64_t)v, lane >= delta ? lane - delta : lane);
    return r;
}
KV_DEV uint64_t shfl_xor64(uint64_t v, int m, int lane) { return kvemu::collective_shfl(v, lane ^ m); }
KV_DEV int shfl_xor32(int v, int m, int lane) { return (int)(int64_t)kvemu::collective_shfl((uint64_t)(int64_t)v, lane ^ m); }
KV_DEV float shfl_xorf(float v, int m, int lane) { return shflf(v, lane ^ m); }
// lanes run one at a time in the emulator, so plain read-modify-write is atomic
KV_DEV uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
KV_DEV void atomic_add_u64(uint64_t* p, uint64_t v) { *p += v; }
KV_DEV uint64_t ldg64(const uint64_t* p) { return *p; }
KV_DEV uint64_t atomic_cas_u64(uint64_t* p, uint64_t expect, uint64_t val) { uint64_t o = *p; if (o == expect) *p = val; return o; }
KV_DEV void atomic_store_u64(uint64_t* p, uint64_t v) { *p = v; }
KV_DEV uint64_t ld_cg_u64(const uint64_t* p) { return *p; }
KV_DEV uint32_t ld_cg_u32(const uint32_t* p) { return *p; }
KV_DEV float ld_cg_f32(const float* p) { return *p; }
KV_DEV void mem_fence() {}
}  // namespace kv
#else
#define KV_DEV __device__ __forceinline__
#define KV_DEVFN __device__
namespace kv {
constexpr unsigned FULL = 0xffffffffu;
KV_DEV int popc32(uint32_t x) { return __popc(x); }
KV_DEV int popc64(uint64_t x) { return __popcll(x); }
KV_DEV int ctz64(uint64_t x) { return __ffsll((long long)x) - 1; }
KV_DEV int msb64(uint64_t x) { return 63 - __clzll((long long)x); }
KV_DEV int ffs32(uint32_t x) { return __ffs((int)x); }
KV_DEV uint64_t shfl64(uint64_t v, int src) { return __shfl_sync(FULL, v, src); }
KV_DEV int shfl32(int v, int src) { return __shfl_sync(FULL, v, src); }
KV_DEV float shflf(float v, int src) { return __shfl_sync(FULL, v, src); }
KV_DEV uint32_t ballot(bool p) { return __ballot_sync(FULL, p); }
KV_DEV void syncwarp() { __syncwarp(); }
KV_DEV int shfl_up32(int v, int delta, int /*lane*/) { return __shfl_up_sync(FULL, v, delta); }
KV_DEV uint64_t shfl_xor64(uint64_t v, int m, int /*lane*/) { return __shfl_xor_sync(FULL, v, m); }
KV_DEV int shfl_xor32(int v, int m, int /*lane*/) { return __shfl_xor_sync(FULL, v, m); }
KV_DEV float shfl_xorf(float v, int m, int /*lane*/) { return __shfl_xor_sync(FULL, v, m); }
KV_DEV uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
KV_DEV void atomic_add_u64(uint64_t* p, uint64_t v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
KV_DEV uint64_t ldg64(const uint64_t* p) { return __ldg(p); }
KV_DEV uint64_t atomic_cas_u64(uint64_t* p, uint64_t expect, uint64_t val) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)expect, (unsigned long long)val);
}
KV_DEV void atomic_store_u64(uint64_t* p, uint64_t v) {
    atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
// L2-coherent loads: data other SMs wrote during the same kernel must not come from this SM's L1
KV_DEV uint64_t ld_cg_u64(const uint64_t* p) { return __ldcg(p); }
KV_DEV uint32_t ld_cg_u32(const uint32_t* p) { return __ldcg(p); }
KV_DEV float ld_cg_f32(const float* p) { return __ldcg(p); }
KV_DEV void mem_fence() { __threadfence(); }
}  // namespace kv
#endif

namespace kv {
// ---- sub-warp collectives: a warp holds 32 / W boards, W consecutive lanes each (W = 16 or 32).  Built on the
// full-warp primitives above, so every lane of the warp must reach them together; q = lane & (W - 1) is the lane's
// index inside its board's group.  xor / up shuffles with distances < W never leave the group.
template <int W> KV_DEV int sub_q(int lane) { return lane & (W - 1); }
template <int W> KV_DEV uint64_t sub_shfl64(uint64_t v, int src, int lane) { return shfl64(v, (lane & ~(W - 1)) | src); }
template <int W> KV_DEV int sub_shfl32(int v, int src, int lane) { return shfl32(v, (lane & ~(W - 1)) | src); }
template <int W> KV_DEV uint32_t sub_ballot(bool p, int lane) {
    const uint32_t b = ballot(p);
    return W == 32 ? b : ((b >> (lane & ~(W - 1) & 31)) & ((W == 32) ? 0xFFFFFFFFu : ((1u << (W & 31)) - 1u)));
}
template <int W> KV_DEV uint64_t sub_or64(uint64_t v, int lane) {
#pragma unroll
    for (int m = W / 2; m >= 1; m >>= 1) v |= shfl_xor64(v, m, lane);
    return v;
}
template <int W> KV_DEV uint64_t sub_sum64(uint64_t v, int lane) {
#pragma unroll
    for (int m = W / 2; m >= 1; m >>= 1) v += shfl_xor64(v, m, lane);
    return v;
}
template <int W> KV_DEV int sub_sum32(int v, int lane) {
#pragma unroll
    for (int m = W / 2; m >= 1; m >>= 1) v += shfl_xor32(v, m, lane);
    return v;
}
template <int W> KV_DEV int sub_max32(int v, int lane) {
#pragma unroll
    for (int m = W / 2; m >= 1; m >>= 1) {
        const int o = shfl_xor32(v, m, lane);
        v = o > v ? o : v;
    }
    return v;
}
template <int W> KV_DEV int sub_incl_scan(int v, int lane) {
    const int q = lane & (W - 1);
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const int t = shfl_up32(v, d, lane);
        if (q >= d) v += t;
    }
    return v;
}
// index of the i-th (0-based) set bit of m (m has more than i bits set): popcount binary search
KV_DEV int nth_set64(uint64_t m, int i) {
    uint32_t x = (uint32_t)m;
    int pos = 0;
    const int c = popc32(x);
    if (i >= c) {
        i -= c;
        x = (uint32_t)(m >> 32);
        pos = 32;
    }
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) {
        const uint32_t low = x & ((1u << sh) - 1u);
        const int cl = popc32(low);
        if (i >= cl) {
            i -= cl;
            x >>= sh;
            pos += sh;
        } else {
            x = low;
        }
    }
    return pos;
}

// inclusive prefix sum over the warp
KV_DEV int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = shfl_up32(v, d, lane);
        if (lane >= d) v += t;
    }
    return v;
}
KV_DEV uint64_t warp_or64(uint64_t v, int lane) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v |= shfl_xor64(v, m, lane);
    return v;
}
KV_DEV uint64_t warp_sum64(uint64_t v, int lane) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor64(v, m, lane);
    return v;
}
KV_DEV int warp_sum32(int v, int lane) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += shfl_xor32(v, m, lane);
    return v;
}
}  // namespace kv
