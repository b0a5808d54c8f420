/*
 * kv_b200.h — C ABI of libkv_b200.so, the B200 (sm_100a) self-play hot path for KnightVision.
 *
 * The reference (TheRealShamsaba/KnightVision) is pure Python and has no FFI; its callers bind to four
 * Python objects.  This header is the boundary a maintainer would bind with ctypes/cffi
 * (see INTEGRATION.md); each entry point names the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; all `d_*` pointers are device pointers on the ctx's
 * GPU (e.g. torch tensor .data_ptr()), all `h_*` pointers are host pointers; `stream` is a cudaStream_t
 * (NULL = default stream).  Every function returns 0 on success, <0 on error (kv_last_error()).
 * There is no CPU fallback: without a CUDA device kv_create fails.
 *
 * Board line (128 B = 16 little-endian u64, one cache line, lane i of the owning warp loads word i):
 *   w[0..11]  bitboards in ai/ai.py:7-10 order  wK wQ wR wB wN wp bK bQ bR bB bN bp;
 *             bit = row*8+col, row 0 = rank 8 (board[row][col], core/chessEngine.py:39-47)
 *   w[12]     bit0 whiteToMove | bits1-6 moved flags wK,bK,wRk,wRq,bRk,bRq (:66-71) | bits8-14
 *             enPassantPossible (64 = none, :72) | bits16-21 whiteKingLocation | bits24-29
 *             blackKingLocation (:59-60) | bits32-47 halfMoveClock (:79)
 *   w[13..15] owned by the engine (perft root id / path hash, MCTS bookkeeping); 0 on input
 * Move word (u16): from | to<<6 | isEnPassantMove<<12 | isCastleMove<<13 | isPawnPromotion<<14
 * (Move, core/chessEngine.py:683-733).  Policy index = from*64+to (ai/ai.py:51-57).
 */
#ifndef KV_B200_H
#define KV_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kv_ctx kv_ctx;

#if defined(__GNUC__)
#define KV_API __attribute__((visibility("default")))
#else
#define KV_API
#endif

#define KV_LINE_WORDS 16
#define KV_MAX_MOVES 256
/* result flags of kv_movegen: checkMate, staleMate, draw50 (core/chessEngine.py:632-651), E3 = the
 * inCheck returned by checkForPinsAndChecks (:325-383), only-kings = isDraw() (:21-33), state-mutated =
 * getKingMoves' restore quirk rewrote the board (:564), overflow = more than KV_MAX_MOVES moves */
#define KV_RF_CHECKMATE 1
#define KV_RF_STALEMATE 2
#define KV_RF_DRAW50 4
#define KV_RF_E3_CHECK 8
#define KV_RF_ONLY_KINGS 16
#define KV_RF_STATE_MUTATED 32
#define KV_RF_OVERFLOW 64

/* ---- context ------------------------------------------------------------------------------------ */
KV_API int kv_create(int device, kv_ctx** out);
KV_API void kv_destroy(kv_ctx* ctx);
KV_API const char* kv_last_error(kv_ctx* ctx);       /* ctx may be NULL: error of the last failed kv_create */
KV_API int kv_abi_version(void);
KV_API int kv_sm_count(kv_ctx* ctx);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
KV_API uint64_t kv_launch_count(kv_ctx* ctx);

/* per-kernel device timing with CUDA events on the launching stream (bench.py's roofline numbers).
 * kv_profile_read synchronises and returns, per kernel category (KV_K_*), the summed duration in ms and
 * the launch count since the previous read; returns the number of categories. */
#define KV_K_MOVEGEN 0
#define KV_K_MAKE_MOVES 1
#define KV_K_PERFT_EXPAND 2
#define KV_K_PERFT_LEAF 3
#define KV_K_ENCODE 4
#define KV_K_NET_STEM 5
#define KV_K_NET_CONV 6
#define KV_K_NET_HEAD 7
#define KV_K_MCTS_SELECT 8
#define KV_K_MCTS_EXPAND 9
#define KV_K_MCTS_MISC 10
#define KV_K_COUNT 11
KV_API int kv_profile_enable(kv_ctx* ctx, int on);
KV_API int kv_profile_read(kv_ctx* ctx, double* ms, uint64_t* n, int cap);

/* ---- rules (replaces GameState.getValidMoves / makeMove, core/chessEngine.py:277-321, :127-197) --- */
/* d_lines [n][16] (rewritten only where KV_RF_STATE_MUTATED), d_moves [n][stride] u16 (stride even,
 * entries >= count unspecified), d_counts [n], d_flags [n] */
KV_API int kv_movegen(kv_ctx* ctx, uint64_t* d_lines, int n, uint16_t* d_moves, int stride, int32_t* d_counts,
               int32_t* d_flags, void* stream);
/* in place; d_moves [n] one move word per board, 0xFFFF = leave the board untouched */
KV_API int kv_make_moves(kv_ctx* ctx, uint64_t* d_lines, int n, const uint16_t* d_moves, void* stream);
/* squareUnderAttack(r,c) (core/chessEngine.py:400-415) for all 64 squares: bit r*8+c of d_masks[i]; inCheck()
 * (:388-394) is the bit of the side to move's king location */
KV_API int kv_attacked(kv_ctx* ctx, const uint64_t* d_lines, int n, uint64_t* d_masks, void* stream);
KV_API int kv_attacked_host(kv_ctx* ctx, const uint64_t* h_lines, int n, uint64_t* h_masks);
/* host-buffer forms (copies inside): what a ctypes binding of GameState would call */
KV_API int kv_movegen_host(kv_ctx* ctx, uint64_t* h_lines, int n, uint16_t* h_moves, int stride, int32_t* h_counts,
                    int32_t* h_flags);
KV_API int kv_make_moves_host(kv_ctx* ctx, uint64_t* h_lines, int n, const uint16_t* h_moves);

/* perft over getValidMoves/makeMove with bulk counting at depth 1 (no reference counterpart: the golden
 * counts of SURVEY.md §8c were produced by exactly this driver over the unmodified engine).
 * d_roots [n][16]; out [n][8] u64: nodes, captures, e.p., castles, promotions (of the leaf moves),
 * order digest (order-sensitive hash of every move list visited), movegen calls, 0.
 * |chunk| = boards per launch (0 = 65536); chunk < 0 = counts only (the digest stays 0; leaves are bulk-counted from
 * the destination sets without laying the ordered lists out). */
KV_API int kv_perft(kv_ctx* ctx, const uint64_t* d_roots, int n, int depth, uint64_t* d_out, int chunk, void* stream);
KV_API int kv_perft_host(kv_ctx* ctx, const uint64_t* h_roots, int n, int depth, uint64_t* h_out, int chunk);

/* ---- encoders (replaces encode_board, ai/ai.py:17-41) ---------------------------------------------- */
/* d_planes [n][12][8][8] float32 one-hot, the reference's record/input format */
KV_API int kv_encode(kv_ctx* ctx, const uint64_t* d_lines, int n, float* d_planes, void* stream);

/* ---- policy/value network (replaces ChessNet, ai/model.py:27-77) ----------------------------------- */
/* Architecture: conv 12->stem_channels, [conv stem->tower if has_conv2], n_blocks x ResidualBlock(tower),
 * heads as in the reference.  The reference net is (256, 512, 5, 1).  max_boards = largest batch. */
KV_API int kv_net_create(kv_ctx* ctx, int stem_channels, int tower_channels, int n_blocks, int has_conv2, int max_boards);
/* Weights: the fp32 tensors of ChessNet.state_dict() concatenated in key order (every *.num_batches_tracked
 * skipped): conv w,b then bn weight,bias,running_mean,running_var for each layer; policy_fc / value_fc as stored.
 * BatchNorm is folded (eval mode, eps 1e-5) and the tower weights are converted to bf16 on the device. */
/* tower kernel variant: 1 = one CTA per 128x256 tile (tcgen05 cta_group::1), 2 = CTA pairs sharing the weight tile
 * (cta_group::2, 256x256 per pair; default).  Bit-identical outputs. */
KV_API int kv_net_set_conv_mode(kv_ctx* ctx, int cta_group);
/* Tower schedule (CTA-pair kernel only).  0 = one launch per layer.  1 = the tower convolutions of a forward pass as ONE
 * launch scheduled by tile dependencies (a 3x3 convolution is board-local: no grid-wide barrier between layers, no
 * per-layer tail); bit-identical to 0.  2 (default) = the whole-tower launch with the halo activation operand: each
 * activation tile is fetched 3 times per 64-channel block instead of 9 (row shifts are descriptor offsets), 11 % faster;
 * the taps are accumulated in another order, so it equals 0 / 1 to fp32 rounding (same 2e-2 tolerance against the fp32
 * reference). */
KV_API int kv_net_set_tower_fused(kv_ctx* ctx, int on);
/* modes 3 / 4 (opt-in): mode 2 on clusters of four CTAs — two CTA pairs on neighbouring board tiles share every weight
 * tile by TMA multicast; bit-identical to mode 2.  Mode 3 uses only the SMs such clusters can cover (132 of 148 on B200:
 * slower); mode 4 also runs the CTA-pair kernel on the remaining SMs, concurrently (+1 % over mode 2).
 * kv_net_tower_clusters4 = how many 4-CTA clusters are co-resident on this GPU (< 8: modes 3 / 4 unavailable). */
KV_API int kv_net_tower_clusters4(kv_ctx* ctx);
KV_API uint64_t kv_net_blob_floats(kv_ctx* ctx);
KV_API int kv_net_load(kv_ctx* ctx, const float* h_blob, uint64_t n_floats);
/* Device staging buffer of kv_net_blob_floats() floats: write the blob there (e.g. as the target of an NCCL
 * broadcast) and call kv_net_commit_weights. */
KV_API void* kv_net_blob_device_ptr(kv_ctx* ctx);
KV_API int kv_net_commit_weights(kv_ctx* ctx, void* stream);
/* The FOLDED weights the kernels read (tower bf16 with BatchNorm folded, biases, stem table, heads) as ONE device
 * buffer of kv_net_folded_bytes() bytes (52 MB for the reference net: half the fp32 blob).  Weight distribution between
 * generations (the reference re-loads a checkpoint per worker, scripts/self_play.py:54-77): the trainer's rank commits,
 * the buffer is NCCL-broadcast into every other rank's buffer, and those ranks call kv_net_adopt_folded (no fold
 * there; clears the evaluation cache like a commit). */
KV_API uint64_t kv_net_folded_bytes(kv_ctx* ctx);
KV_API void* kv_net_folded_device_ptr(kv_ctx* ctx);
KV_API int kv_net_adopt_folded(kv_ctx* ctx, void* stream);
/* forward(x) -> (policy logits [n][4096] fp32, value [n] fp32 after tanh); either output may be NULL.
 * Input = board lines (the stem kernel fuses encode_board), or the reference's fp32 one-hot planes. */
KV_API int kv_net_forward(kv_ctx* ctx, const uint64_t* d_lines, int n, float* d_policy, float* d_value, void* stream);
/* test hook: stem + the first n_convs tower convolutions; copies the NHWC bf16 activations [n*64][*channels] out */
KV_API int kv_net_forward_partial(kv_ctx* ctx, const uint64_t* d_lines, int n, int n_convs, void* d_act_out, int* channels);
KV_API int kv_net_forward_planes(kv_ctx* ctx, const float* d_planes, int n, float* d_policy, float* d_value, void* stream);

/* ---- self-play engine: batched PUCT search + game records (replaces the game loop of scripts/self_play.py:111-255;
 *      the tree search itself is new functionality, specified in DESIGN.md §MCTS and oracle/kv_oracle.c) ----------- */
/* n_games concurrent games on this GPU, `sims` simulations per move (one in flight per game), edges_per_node = edge
 * pool sizing (0 = 64), max_plies = ply cap (draw), temp_plies = plies that sample the move from the visit counts,
 * eval_mode 0 = hash test evaluator, 1 = the network of kv_net_create (max_boards >= n_games). */
KV_API int kv_mcts_create(kv_ctx* ctx, int n_games, int sims, int edges_per_node, int max_plies, int temp_plies,
                          float c_puct, float dir_alpha, float dir_eps, uint64_t seed, int eval_mode);
/* Same with `inflight` = K simulations in flight per game and wave (kv_mcts_create: K = 1).  K > 1 is for fewer
 * games than the network batch wants (n_games * K leaves per wave; kv_net_create max_boards >= n_games * K): the K
 * selections of a wave see each other through VIRTUAL LOSS (an in-flight simulation counts as one visit and one
 * loss on every edge of its path until it is backed up; north_star item 2 / SURVEY 8a row M3).  Deterministic:
 * selections and backups of a game run in slot order; oracle/kv_oracle.c emulates the same waves bit for bit. */
KV_API int kv_mcts_create_k(kv_ctx* ctx, int n_games, int sims, int edges_per_node, int max_plies, int temp_plies,
                            float c_puct, float dir_alpha, float dir_eps, uint64_t seed, int eval_mode, int inflight);
/* d_start [n_games][16] start lines or NULL (initial position); game ids = game_id_base + local index (RNG keys) */
KV_API int kv_mcts_reset(kv_ctx* ctx, const uint64_t* d_start, uint64_t game_id_base, void* stream);
KV_API int kv_mcts_run_sims(kv_ctx* ctx, int n_waves, void* stream);   /* n_waves waves (K = 1: one simulation each) */
KV_API int kv_mcts_finish_move(kv_ctx* ctx, void* stream);            /* pick + record + play the move, reset trees */
KV_API int kv_mcts_run_move(kv_ctx* ctx, void* stream);               /* waves until `sims` simulations ran + finish_move */
/* Evaluation cache: 2^log2_slots entries x 640 B keyed by the 12 bitboards (the network's whole input; 0 = off).
 * A hit skips the tower; priors are re-derived from the cached policy features, so visit counts and games are
 * bit-identical with the cache on or off.  Cleared automatically by kv_net_commit_weights / kv_net_load. */
KV_API int kv_mcts_enable_cache(kv_ctx* ctx, int log2_slots);
KV_API int kv_mcts_cache_clear(kv_ctx* ctx, void* stream);
/* h_out10: games done, sum of sims done in the current move, tower evaluations, plies played, games with edge-pool
 * overflow, white wins, black wins, draws, expansions served by the cache (hits + in-wave duplicates), games stopped
 * by an illegal scripted move */
KV_API int kv_mcts_status(kv_ctx* ctx, uint64_t* h_out10, void* stream);
/* Game-loop rules of scripts/self_play.py:180-199 applied after every move, in the reference's order: only-kings
 * draw (:180), resignation (:185-189: more than min_plies plies played and the network's value of the position the
 * move was chosen in < threshold => -1 if white is to move, else +1), ply cap (:196), then no-moves at the loop head
 * (:125: checkmate :217 / stalemate :221).  Defaults are the reference's: threshold -0.7, min_plies 15;
 * min_plies < 0 switches resignation off. */
KV_API int kv_mcts_set_resign(kv_ctx* ctx, float threshold, int min_plies);
/* Root priors.  0: softmax over the legal moves' logits, Dirichlet noise over the legal moves (tree search).
 * 1: the reference's rule (scripts/self_play.py:150-167) — softmax over ALL 4096 logits, Dirichlet(alpha) noise over
 * all 4096 indices, (1-eps) p + eps noise, then the legal entries renormalised.  -1 = default: 1 when sims == 1 (no
 * search: the move is sampled from these priors exactly as the reference samples it), else 0. */
KV_API int kv_mcts_set_root_mix(kv_ctx* ctx, int mode);
/* Scripted play (replay of recorded games through the game loop): d_moves [n_games][stride] u16 move words matched on
 * (from, to) against the legal moves (0xFFFF = choose as usual; an illegal one stops the game and is counted in
 * kv_mcts_status), d_values [n_games][stride] the value the resign rule sees at that ply (NaN = the evaluator's).
 * Either may be NULL; caller-owned device memory; (NULL, NULL, 0) clears. */
KV_API int kv_mcts_set_script(kv_ctx* ctx, const uint16_t* d_moves, const float* d_values, int stride);
KV_API int kv_mcts_get_roots(kv_ctx* ctx, uint64_t* d_lines, void* stream); /* current position of every game [n][16] */
KV_API int kv_mcts_geometry(kv_ctx* ctx, int32_t* out4);              /* n_games, node_cap, edge_cap, rec_cap */
KV_API int64_t kv_mcts_waves(kv_ctx* ctx);                            /* search waves launched since create */
/* Pipelined search (opt-in): the games are split into two groups whose waves alternate on two CUDA streams (the
 * caller's and an internal one, forked and joined inside every kv_mcts_run_* call) plus a high-priority stream for the
 * tensor-core kernels, so that the tree kernels of one group run on the CUDA cores while the tower of the other group
 * runs.  mode -1 = default (off, unless the environment sets KV_MCTS_PIPELINE=1), 0 = off, 1 = on.  Search results are
 * bit-identical either way (games are independent; the shared evaluation cache is transparent).  Measured gain on
 * B200: about 1 % — the step is bound by the power cap, not by idle SMs. */
KV_API int kv_mcts_set_pipeline(kv_ctx* ctx, int mode);
/* Evaluator schedule after the tower: 0 (default) = one kernel, one CTA per leaf; 1 = a streaming head-features kernel
 * + a kernel that finishes eight leaves per CTA (value MLP for all of them from one pass over value_fc1, then a warp
 * per leaf for the legal-move logits, priors and backup); -1 = default (KV_MCTS_EVAL_SPLIT).  Bit-identical results.
 * Measured on B200: the evaluator gets 10 % shorter and the power-capped step does not change. */
KV_API int kv_mcts_set_eval_split(kv_ctx* ctx, int mode);
/* records of every game in game order; d_lines [cap][16] board lines (kv_encode gives the reference's planes),
 * d_move policy index (ai/ai.py:51-57), d_reward 1.0 / 0.2 / -1.0 (scripts/self_play.py:245-250), d_game index */
KV_API int kv_mcts_records(kv_ctx* ctx, uint64_t* d_lines, int32_t* d_move, float* d_reward, int32_t* d_game, int cap,
                           int32_t* h_count, void* stream);
/* test hooks (host buffers): root edges of one game; per-node evaluator values / per-edge priors for oracle replay */
KV_API int kv_mcts_read_root(kv_ctx* ctx, int game, uint16_t* h_moves, uint32_t* h_N, float* h_W, float* h_P, int32_t* h_info4);
KV_API int kv_mcts_dump_tree(kv_ctx* ctx, int game, float* h_node_val, int32_t* h_node_first, float* h_edge_P,
                             uint64_t* h_root_line16);

/* ---- training-side convolution operators (the tower convolutions of ai/model.py:19-25,58-59 as
 *      scripts/train.py:158-181 runs them forward and backward under autocast; cuDNN in the reference) -------------
 * All tensors are caller-owned device memory: activations NHWC bf16 [n_boards][8][8][C] (torch channels_last), the
 * parameter / its gradient fp32 [Cout][Cin][3][3] as PyTorch stores them.  cin % 64 == 0, cout % 256 == 0, <= 512.
 *   kv_conv3x3_pack   parameter -> the kernels' bf16 [Cout][9][Cin]; flip_transpose = 1 gives the dgrad operand
 *                     [Cin][9][Cout] with the taps mirrored
 *   kv_conv3x3_fprop  y = [relu](conv3x3(x, w) + bias [+ residual]) (bias / residual may be NULL).  dgrad is the same
 *                     call on (dy, flip-transposed pack, cin <-> cout)
 *   kv_conv3x3_wgrad  dw[co][ci][ky][kx] = sum_{b,y,x} dy[b,y,x,co] * x[b,y+ky-1,x+kx-1,ci], fp32, deterministic
 *                     (cin % 256 == 0, cout % 128 == 0) */
KV_API int kv_conv3x3_pack(kv_ctx* ctx, const float* d_w, int cout, int cin, int flip_transpose, void* d_out_bf16,
                           void* stream);
KV_API int kv_conv3x3_fprop(kv_ctx* ctx, const void* d_x, const void* d_w_packed, const float* d_bias,
                            const void* d_residual, void* d_y, int n_boards, int cin, int cout, int relu, void* stream);
KV_API int kv_conv3x3_wgrad(kv_ctx* ctx, const void* d_x, const void* d_dy, float* d_dw, int n_boards, int cin, int cout,
                            void* stream);

/* Train-mode BatchNorm2d + ReLU (+ residual add) on NHWC bf16 [rows = boards * 64][C] (ai/model.py:19-25,58-59 under
 * scripts/train.py:158-181; torch.nn.BatchNorm2d semantics: batch statistics with biased variance, running statistics
 * updated with `momentum` and the unbiased variance, eps inside the square root).  C in {64,128,256,512,1024}.
 *   fwd  y = [relu](gamma * (z - mean) * rstd + beta [+ residual]); writes save_mean / save_rstd [C] for the backward;
 *        running_mean / running_var may be NULL
 *   bwd  g = dy * [y > 0]; dgamma = sum g * zhat; dbeta = sum g; dz = gamma * rstd * (g - dbeta/N - zhat * dgamma/N);
 *        d_dres (nullable) receives g, the gradient of the residual input.  Reductions are deterministic.
 *   kv_channel_sum  out[c] = sum over rows of x[r][c] (the convolution-bias gradient) */
KV_API int kv_bn_relu_fwd(kv_ctx* ctx, const void* d_z, const void* d_residual, const float* d_gamma, const float* d_beta,
                          float* d_running_mean, float* d_running_var, float momentum, float eps, void* d_y,
                          float* d_save_mean, float* d_save_rstd, int rows, int C, int relu, void* stream);
KV_API int kv_bn_relu_bwd(kv_ctx* ctx, const void* d_dy, const void* d_y, const void* d_z, const float* d_gamma,
                          const float* d_save_mean, const float* d_save_rstd, void* d_dz, void* d_dres, float* d_dgamma,
                          float* d_dbeta, int rows, int C, int relu, void* stream);
KV_API int kv_channel_sum(kv_ctx* ctx, const void* d_x, float* d_out, int rows, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KV_B200_H */
