// kv_net.h — device-resident network state (kv_net.cu), shared with the MCTS translation unit
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <vector>

struct kv_conv {
    __nv_bfloat16* w = nullptr;   // [cout][9][cin], BN scale folded
    float* b = nullptr;           // [cout]
    int cin = 0, cout = 0;
    CUtensorMap map;        // 2-D boxes {64, 256}: the whole N tile (1-CTA kernel)
    CUtensorMap map_half;   // 2-D boxes {64, 128}: this CTA's half of the N tile (2-CTA kernel)
    CUtensorMap map_q;      // 2-D boxes {64, 64}: a quarter of the N tile (4-CTA clusters, multicast)
};

struct kv_net {
    int C1 = 256, C = 512, blocks = 5;
    bool has_conv2 = true;
    bool loaded = false;
    int cap = 0;                           // boards (even)
    std::vector<kv_conv> convs;
    __nv_bfloat16* act[3] = {nullptr, nullptr, nullptr};   // NHWC bf16 [cap*64][max(C1,C)]
    CUtensorMap map_act[3][2];             // [buffer][0: C1-channel view, 1: C-channel view]
    CUtensorMap map_act_h[3][2];           // same views, dims (c, x, board, y) and boxes {64, 8, 2, 10}: the halo variant
    float *stem_table = nullptr, *stem_bias = nullptr;
    float *wh = nullptr, *bh = nullptr;    // head 1x1 convs [3][C], [3]
    float *wfc = nullptr, *bfc = nullptr;  // policy_fc [4096][128], [4096]
    float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;   // value_fc1 [512][64], value_fc2 [512]
    void* d_folded = nullptr;              // arena holding every folded weight above (one NCCL broadcast moves a generation)
    size_t folded_bytes = 0;
    float* d_blob = nullptr;               // fp32 state_dict staging (NCCL broadcast target)
    size_t blob_floats = 0;
    int conv_mode = 2;                     // 1: cta_group::1 kernel, 2: cta_group::2 CTA-pair kernel
    int tower_fused = 2;                   // conv_mode 2 only: 1 = the whole tower as one dependency-scheduled launch,
                                           // 2 (default) = the same with the halo A-operand (3 instead of 9 fetches),
                                           // 3 = halo operand + 4-CTA clusters sharing the weight tiles by multicast
    void* d_layers = nullptr;              // TowerLayerDev[convs.size()] (kv_net.cu)
    uint32_t* d_done[2] = {nullptr, nullptr};   // tile-completion counters [layers][m_stride], one set per game group
    int m_stride = 0;
    bool halo_ok = true;                   // the halo tensor maps could be encoded
    cudaStream_t side = nullptr;           // hybrid tower launch: the CTA-pair kernel's stream, forked / joined per launch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int clusters4 = 0;                     // co-resident 4-CTA clusters of tower_umma4_kernel (0: variant unavailable)
    int tower_chunk = 74;                  // board tiles per depth-first chunk of the whole-tower launch (0: layer-major)
    int* d_flag = nullptr;
    uint64_t* d_lines_tmp = nullptr;
};

struct kv_ctx;
// n = boards (grid sizing / upper bound); n_ptr = optional device-side count that overrides n inside the kernels;
// board_base = first board of this launch inside the activation buffers (the result is at act[*final_buf] +
// board_base * 64 * max(C, C1)); conv_stream + handoff = run the tensor-core kernels on another stream (the stem stays
// on st; handoff is recorded on st and waited for by conv_stream; the caller orders st after conv_stream again)
int kv_net_tower(kv_ctx* ctx, const uint64_t* d_lines, int n, cudaStream_t st, int* final_buf, int max_convs = -1,
                 const int* n_ptr = nullptr, int board_base = 0, cudaStream_t conv_stream = nullptr,
                 cudaEvent_t handoff = nullptr, int stem_grid = 0);   // stem_grid > 0: CTAs of the (grid-stride) stem

// training path (kv_train.cu): activation tensor map over a caller-owned NHWC bf16 tensor [boards][8][8][C] with TMA
// boxes of {64 channels, 8, 8, box_boards}; one 3x3 convolution through the tower kernel on caller-owned tensors
int kv_make_act_map(kv_ctx* ctx, CUtensorMap* m, void* base, int C, int boards, int box_boards, int pitch = 0);
int kv_conv_launch(kv_ctx* ctx, const __nv_bfloat16* x, const __nv_bfloat16* w_packed, const float* bias,
                   const __nv_bfloat16* residual, __nv_bfloat16* y, int n, int cin, int cout, int relu, cudaStream_t st);
