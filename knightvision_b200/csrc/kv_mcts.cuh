// kv_mcts.cuh — batched PUCT tree search over a GPU-resident node pool: warp-per-game device code.
//
// The reference has no tree search; the algorithm is specified in DESIGN.md §MCTS and restated
// sequentially in oracle/kv_oracle.c (mcts_search), which these functions must match bit-for-bit.
// cfg.inflight = K simulations are in flight per game and wave.  K = 1 (4 096 concurrent games already fill the
// network batch) makes the search inside a game sequential; K > 1 (fewer games per GPU) selects K leaves per wave
// with VIRTUAL LOSS: every edge / node on an in-flight path carries one virtual visit (kept in the top 8 bits of
// the visit counters, so it never mixes with the fp32 W sums) that counts as a loss in the PUCT score until the
// simulation is backed up.  The warp that owns the game performs the K selections one after the other and the
// backups in slot order, so the result is deterministic and bit-exact with the oracle's wave emulation.
// A warp owns a game:
//   select   warp-level argmax of the PUCT score over a node's edges (lane k handles edges k, k+32, ...),
//            ties broken towards the lowest edge index = the reference's move order
//   expand   make-move + legal move generation for the new leaf (kv_rules.cuh), edges appended to the game's pool
//   evaluate the leaf is queued (one atomic add) for the batched network pass; terminal leaves back up at once
//   backup   lane 0 walks the recorded path: W += +/-v, N += 1 in the oracle's order (fp32, no atomics needed)
// All floating point that steers the search goes through include/kv_detmath.h (no FMA, fixed order).
// Compiles for the device and, with KV_HOST_EMU, for the CPU lock-step emulator (tests/simt_emu).
#pragma once
#include "../../include/kv_detmath.h"
#include "kv_rules.cuh"

namespace kv {

struct MctsCfg {
    int sims, node_cap, edge_cap, temp_plies, max_plies, eval_mode;   // eval_mode: 0 hash evaluator, 1 network
    float c_puct, dir_alpha, dir_eps;
    int rec_cap;          // records per game (>= max_plies)
    uint32_t cache_mask;  // evaluation cache slots - 1 (power of two), 0 = cache disabled
    int inflight;         // K: simulations in flight per game and wave (>= 1); slot index gs = game * K + j
    uint64_t seed;
    // resignation (scripts/self_play.py:184-189): after a move, once MORE than resign_min_plies plies were played and the
    // network's value of the position the move was chosen in (raw output, no perspective flip) is < resign_thr, the game
    // ends with result -1 if white is to move, else +1.  resign_min_plies < 0: never.  Reference: -0.7 / 15.
    float resign_thr;
    int resign_min_plies;
    // root priors: 0 = softmax over the legal moves, Dirichlet noise over the legal moves (search default);
    // 1 = the reference's rule (scripts/self_play.py:150-167): softmax over ALL 4096 indices, Dirichlet noise over all
    //     4096 indices, mix, then renormalise over the legal moves
    int root_mix;
    int script_stride;    // plies per game of the script arrays (0: no script)
};
constexpr int POLICY_N = 4096;   // policy indices (ai/ai.py:51-57)

// visit counters: low 24 bits real visits, top 8 bits virtual visits of in-flight simulations (0 between waves)
constexpr uint32_t VL_ONE = 1u << 24;
constexpr uint32_t N_MASK = VL_ONE - 1u;
constexpr int NODE_TERM = 1 << 16;      // NodeMeta.ne_term: terminal node
constexpr int NODE_PENDING = 1 << 17;   // expanded this wave, priors not delivered yet

// Evaluation cache entry: the network's input is the 12 bitboards only (ai/ai.py:17-41 has no side/castling/e.p.
// planes), so the key is the bitboards and the payload is what the heads need to produce ANY legal-move logit:
// the 128 policy features + the value.  A hit skips the tower; priors are re-derived exactly as on a miss, so
// search results are bit-identical with the cache on or off.
struct alignas(128) CacheEntry {   // 640 B
    uint64_t tag;      // 0 empty, ~0 locked (being rewritten), else position hash
    uint32_t stamp;    // wave id of the last claim / hit
    uint32_t pend;     // 0: payload valid; 1 + eval slot: under evaluation in wave `stamp`
    uint64_t bb[12];
    float v;           // white-perspective value
    float pad[3];
    float hp[128];     // policy-head features (relu(bn(conv1x1)), flatten order)
};
constexpr int FEAT = 132;   // hp[128], v, 3 pad
constexpr uint64_t CACHE_LOCKED = ~0ull;
constexpr int CACHE_WINDOW = 4;

struct GameHdr {          // 64 B
    int n_nodes, n_edges, ply, done;
    int result, n_pend, pad2, overflow;   // n_pend: simulations of this wave waiting for their evaluation
    int sims_done, n_evals, cache_hits, pad0;
    uint64_t game_id;
    uint64_t pad1;
};

struct NodeMeta {         // 16 B: first_edge | n_edges + (term << 16) | visits N | value (terminal or evaluator's)
    int first_edge;
    int ne_term;
    uint32_t N;
    float val;
};

struct MctsArrays {
    GameHdr* hdr;          // [G]
    uint64_t* root_line;   // [G][16]
    uint64_t* node_line;   // [G][node_cap][16]
    NodeMeta* node_meta;   // [G][node_cap]
    float* eP;             // [G][edge_cap]
    uint32_t* eN;
    float* eW;
    int* eChild;
    uint16_t* eMv;
    int* path_edge;        // [G*K][node_cap + 1] (edge index inside the game's pool), one path per in-flight slot
    int* path_node;
    int* pend_node;        // [G*K] leaf waiting for its evaluation (-1 none)
    int* pend_depth;       // [G*K] length of its path
    int* pend_kind;        // [G*K] 1 evaluated by the tower, 2 served by the cache (set by expand, read by commit)
    // evaluation queue of the current wave
    uint32_t* n_eval;      // device counter
    int* eval_game;        // [G*K] slot index gs of the queued leaf
    uint64_t* eval_lines;  // [G*K][16]
    // evaluation cache + "late" queue (cache hits and in-wave followers: expanded after the network pass)
    CacheEntry* cache;     // [cache_mask + 1]
    int* eval_centry;      // [G*K] cache entry claimed for the eval slot, -1 none
    uint64_t* eval_hash;   // [G*K]
    uint32_t* n_late;      // device counter
    int* late_game;        // [G*K] slot index gs
    int* late_src;         // [G*K] -1: features already in feat_game[gs]; else eval slot of the leader
    float* feat_game;      // [G*K][FEAT]
    float* feat_slot;      // [G*K][FEAT] features of every evaluated slot of this wave
    // game records (scripts/self_play.py:173-174): position bitboards + the move played
    uint64_t* rec_line;    // [G][rec_cap][12]
    uint16_t* rec_move;    // [G][rec_cap]
    // Pipelined search (kv_mcts.cu): the games are split into two groups whose waves alternate on two streams.  A
    // launch gets a VIEW of these arrays: n_eval / n_late / eval_* / late_* point at the group's own queues, while
    // feat_slot stays global and is indexed by slot_base + (slot inside the group's queue) — the index a cache entry
    // under evaluation publishes, so that a leaf of one group can follow a leader of the other group.
    int slot_base = 0;
    uint32_t peer_wave = 0;   // wave id of the other group's wave still in flight (0 = none)
    // Scripted play (kv_mcts_set_script): replay of recorded games through the engine's game loop.  script_move
    // [G][script_stride] move words (0xFFFF: choose as usual), matched on (from, to) against the root's legal moves;
    // script_val [G][script_stride] the value the resign rule sees at that ply (NaN: the evaluator's).
    const uint16_t* script_move = nullptr;
    const float* script_val = nullptr;
};
constexpr int HDR_OVERFLOW = 1;      // GameHdr.overflow bits: edge pool overflow
constexpr int HDR_SCRIPT_MISS = 2;   // a scripted move was not legal in its position (game stopped as a draw)

KV_DEV float f_from_bits(uint32_t u) { return kvd_u2f(u); }

KV_DEV uint64_t pos_hash_warp(uint64_t w, int lane) {
    // same fold as the oracle's pos_hash: sequential over the 12 bitboards (cheap: 12 steps)
    uint64_t h = 0x243F6A8885A308D3ull;
#pragma unroll
    for (int p = 0; p < 12; p++) h = kvd_mix64(h ^ (shfl64(w, p) + (uint64_t)(p + 1) * 0x9E3779B97F4A7C15ull));
    return h;
}
KV_DEV float hash_logit(uint64_t ph, int idx) {
    return (float)kvd_rand24(ph, (uint64_t)idx, 1, 0) * (4.0f / 16777216.0f) - 2.0f;
}
KV_DEV float hash_value(uint64_t ph) { return (float)kvd_rand24(ph, 4096, 2, 0) * (2.0f / 16777216.0f) - 1.0f; }
KV_DEV int move_index(int mv) { return (mv & 63) * 64 + ((mv >> 6) & 63); }   // encode_move, ai/ai.py:51-57

// Backup of one simulation, all lanes: lane i owns path level i (i + 32, ...).  The value alternates sign per level
// (exact), and a path never holds the same edge or node twice, so the levels are independent: W += +/-v, the virtual
// visit taken at selection becomes the real one.  Same bits as walking the path sequentially.
KV_DEV void mcts_backup_warp(int lane, const MctsArrays& A, size_t ebase, size_t nbase, size_t pbase, int depth, float v) {
    for (int i = lane; i < depth; i += 32) {
        const float vi = ((depth - i) & 1) ? -v : v;
        const size_t e = ebase + (size_t)A.path_edge[pbase + i];
        A.eW[e] = A.eW[e] + vi;
        A.eN[e] = A.eN[e] + 1u - VL_ONE;
        A.node_meta[nbase + (size_t)A.path_node[pbase + i]].N += 1u - VL_ONE;
    }
    syncwarp();
}
// A selection that ran into a leaf still waiting for its evaluation gives its virtual visits back
KV_DEV void mcts_unwind_warp(int lane, const MctsArrays& A, size_t ebase, size_t nbase, size_t pbase, int depth) {
    for (int i = lane; i < depth; i += 32) {
        A.eN[ebase + (size_t)A.path_edge[pbase + i]] -= VL_ONE;
        A.node_meta[nbase + (size_t)A.path_node[pbase + i]].N -= VL_ONE;
    }
    syncwarp();
}
// Finish the pending simulation of slot gs (its leaf's value is in the node): backup + counters.  All lanes.
KV_DEV void mcts_commit_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int g, int gs) {
    GameHdr* h = &A.hdr[g];
    const size_t nbase = (size_t)g * cfg.node_cap, ebase = (size_t)g * cfg.edge_cap;
    const size_t pbase = (size_t)gs * (cfg.node_cap + 1);
    const int c = A.pend_node[gs];
    mcts_backup_warp(lane, A, ebase, nbase, pbase, A.pend_depth[gs], A.node_meta[nbase + c].val);
    if (lane == 0) {
        h->sims_done += 1;
        if (A.pend_kind[gs] == 2) h->cache_hits += 1;
        else h->n_evals += 1;
        A.pend_node[gs] = -1;
    }
    syncwarp();
}

// ---- evaluation cache ------------------------------------------------------------------------------------------
KV_DEV uint64_t cache_hash(uint64_t h) { return (h == 0 || h == CACHE_LOCKED) ? 0x9E3779B97F4A7C15ull : h; }

// Probe the cache for the position in w (lanes 0-11 hold the bitboards).  Returns 1 = hit (features copied to
// feat_game[g]), 2 = the same position is under evaluation in this wave (aux = leader's eval slot), 0 = miss.
// Readers validate with a seqlock on the tag: writers in the same kernel set tag = LOCKED before touching an entry.
KV_DEV int cache_lookup_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, uint32_t wave, int g, uint64_t w,
                             uint64_t h, int& aux) {
    const uint32_t base = (uint32_t)h & cfg.cache_mask;
    aux = -1;
    for (int attempt = 0; attempt < 16; attempt++) {
        uint64_t t = 0;
        if (lane < CACHE_WINDOW) t = ld_cg_u64(&A.cache[(base + lane) & cfg.cache_mask].tag);
        const uint32_t match = ballot(lane < CACHE_WINDOW && t == h);
        const uint32_t locked = ballot(lane < CACHE_WINDOW && t == CACHE_LOCKED);
        if (match) {
            CacheEntry* e = &A.cache[(base + (uint32_t)(ffs32(match) - 1)) & cfg.cache_mask];
            mem_fence();
            const uint64_t kb = lane < 12 ? ld_cg_u64(&e->bb[lane]) : 0ull;
            const bool same = ballot(lane < 12 && kb != w) == 0;
            const uint32_t stamp = ld_cg_u32(&e->stamp), pend = ld_cg_u32(&e->pend);
            mem_fence();   // the other group's evaluator may be filling this entry right now: payload after pend == 0
            float f0 = ld_cg_f32(&e->hp[lane * 4]), f1 = ld_cg_f32(&e->hp[lane * 4 + 1]);
            float f2 = ld_cg_f32(&e->hp[lane * 4 + 2]), f3 = ld_cg_f32(&e->hp[lane * 4 + 3]);
            const float v = ld_cg_f32(&e->v);
            mem_fence();
            const uint64_t t2 = ld_cg_u64(&e->tag);
            const bool stable = ballot(t2 != h) == 0;
            if (stable && same) {
                if (pend == 0) {
                    float* dst = A.feat_game + (size_t)g * FEAT;
                    dst[lane * 4] = f0; dst[lane * 4 + 1] = f1; dst[lane * 4 + 2] = f2; dst[lane * 4 + 3] = f3;
                    if (lane == 0) {
                        dst[128] = v;
                        e->stamp = wave;
                    }
                    return 1;
                }
                if (stamp == wave || (A.peer_wave && stamp == A.peer_wave)) {   // leader's slot in feat_slot (global)
                    aux = (int)pend - 1;
                    return 2;
                }
                return 0;   // stale pending entry (its evaluation was never delivered): plain miss
            }
            if (stable) return 0;   // 64-bit hash collision with a different position: miss
            continue;               // entry changed under us: probe again
        }
        if (!locked) return 0;
        // a window slot is being rewritten right now (possibly with this very position): look again
    }
    return 0;
}

// After a miss: try to claim a window slot for this position, marked "under evaluation by eval slot `slot`".
// Returns the entry index or -1.  Victim = an empty slot, else the least recently stamped one not touched this wave.
KV_DEV int cache_claim_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, uint32_t wave, uint64_t w, uint64_t h,
                            int slot) {
    const uint32_t base = (uint32_t)h & cfg.cache_mask;
    uint64_t t = 0;
    uint32_t key = 0xFFFFFFFFu;
    if (lane < CACHE_WINDOW) {
        const CacheEntry* e = &A.cache[(base + lane) & cfg.cache_mask];
        t = ld_cg_u64(&e->tag);
        const uint32_t stamp = ld_cg_u32(&e->stamp);
        if (t == 0) key = 0;
        else if (t != CACHE_LOCKED && stamp != wave && !(A.peer_wave && stamp == A.peer_wave))
            key = 1u + (stamp & 0x7FFFFFFFu);
    }
    // min key over lanes 0..3 (ties -> lowest lane)
    uint32_t bk = key;
    int bl = lane;
#pragma unroll
    for (int d = 2; d >= 1; d >>= 1) {
        const uint32_t ok = (uint32_t)shfl_xor32((int)bk, d, lane);
        const int ol = shfl_xor32(bl, d, lane);
        if (ok < bk || (ok == bk && ol < bl)) {
            bk = ok;
            bl = ol;
        }
    }
    bk = (uint32_t)shfl32((int)bk, 0);
    bl = shfl32(bl, 0);
    if (bk == 0xFFFFFFFFu) return -1;
    const uint64_t seen = shfl64(t, bl);
    const uint32_t idx = (base + (uint32_t)bl) & cfg.cache_mask;
    CacheEntry* e = &A.cache[idx];
    int ok = 0;
    if (lane == 0) ok = atomic_cas_u64(&e->tag, seen, CACHE_LOCKED) == seen;
    ok = shfl32(ok, 0);
    if (!ok) return -1;
    if (lane < 12) e->bb[lane] = w;
    if (lane == 12) e->stamp = wave;
    if (lane == 13) e->pend = 1u + (uint32_t)(A.slot_base + slot);
    mem_fence();
    syncwarp();
    if (lane == 0) atomic_store_u64(&e->tag, h);
    return (int)idx;
}

// After the evaluation of eval slot `slot`: publish its features for this wave's followers and, if the claimed
// cache entry is still ours, fill its payload.  hp may be null (hash evaluator: only the value is meaningful).
KV_DEV void cache_fill_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, uint32_t wave, int slot, const float* hp,
                            float v_white) {
    float* fs = A.feat_slot + (size_t)(A.slot_base + slot) * FEAT;
    for (int i = lane; i < 128; i += 32) fs[i] = hp ? hp[i] : 0.0f;
    if (lane == 0) fs[128] = v_white;
    const int ci = A.eval_centry[slot];
    if (ci < 0) return;
    CacheEntry* e = &A.cache[ci];
    if (e->tag != A.eval_hash[slot] || e->pend != 1u + (uint32_t)(A.slot_base + slot) || e->stamp != wave) return;
    for (int i = lane; i < 128; i += 32) e->hp[i] = hp ? hp[i] : 0.0f;
    if (lane == 0) e->v = v_white;
    mem_fence();
    syncwarp();
    if (lane == 0) e->pend = 0;
}

// One selection for game g into in-flight slot gs: descend (virtual-loss-aware PUCT), create the leaf, then queue it
// for evaluation or back up a terminal value at once.
// Returns 0 = simulation completed (terminal), 1 = leaf queued (pending), 2 = ran into a pending leaf (nothing done).
KV_DEV int mcts_select_one_warp(const Tables& T, int lane, const MctsCfg& cfg, const MctsArrays& A, int g, int gs,
                                uint16_t* mv, uint32_t wave) {
    GameHdr* h = &A.hdr[g];
    const size_t nbase = (size_t)g * cfg.node_cap, ebase = (size_t)g * cfg.edge_cap;
    const size_t pbase = (size_t)gs * (cfg.node_cap + 1);
    const int n_nodes = h->n_nodes;
    int depth = 0, node = 0, leaf = -1;
    float v = 0.0f;
    uint64_t w = 0;
    if (n_nodes == 0) {
        w = lane < LINE_WORDS ? A.root_line[(size_t)g * LINE_WORDS + lane] : 0ull;
        leaf = 0;
    } else {
        for (;;) {
            const NodeMeta m = A.node_meta[nbase + node];
            const int ne = m.ne_term & 0xFFFF;
            if (m.ne_term & NODE_PENDING) {   // only with K > 1: this leaf's priors arrive at the end of the wave
                syncwarp();
                mcts_unwind_warp(lane, A, ebase, nbase, pbase, depth);
                return 2;
            }
            if (m.ne_term & NODE_TERM) {
                v = m.val;
                if (lane == 0) A.node_meta[nbase + node].N = m.N + 1;
                break;
            }
            // visits seen by the score = real + virtual; an in-flight simulation counts as a loss (W - vl)
            const float sq = KVD_SQRTF((float)((m.N & N_MASK) + (m.N >> 24)));
            float bs = 0.0f;
            int bi = 0x7FFFFFFF;
            for (int k = lane; k < ne; k += 32) {
                const size_t e = ebase + m.first_edge + k;
                const uint32_t en = A.eN[e];
                const float sc = kvd_puct(A.eW[e] - (float)(en >> 24), (en & N_MASK) + (en >> 24), A.eP[e], sq, cfg.c_puct);
                if (bi == 0x7FFFFFFF || sc > bs) {
                    bs = sc;
                    bi = k;
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const float os = shfl_xorf(bs, d, lane);
                const int oi = shfl_xor32(bi, d, lane);
                if (oi != 0x7FFFFFFF && (bi == 0x7FFFFFFF || os > bs || (os == bs && oi < bi))) {
                    bs = os;
                    bi = oi;
                }
            }
            const int ei = m.first_edge + bi;
            if (lane == 0) {
                A.path_node[pbase + depth] = node;
                A.path_edge[pbase + depth] = ei;
                A.eN[ebase + ei] += VL_ONE;               // virtual visit until the backup
                A.node_meta[nbase + node].N = m.N + VL_ONE;
            }
            depth++;
            const int child = A.eChild[ebase + ei];
            if (child < 0) {
                w = lane < LINE_WORDS ? A.node_line[(nbase + node) * LINE_WORDS + lane] : 0ull;
                w = make_move_warp(lane, w, A.eMv[ebase + ei], T_Q);
                leaf = n_nodes;
                if (lane == 0) A.eChild[ebase + ei] = leaf;
                break;
            }
            node = child;
        }
    }
    bool queued = false;
    if (leaf >= 0) {
        const GenOut go = movegen_warp(T, lane, w, mv);   // w carries the :564 rewrite, as the reference's state would
        const int n = go.n < MAX_MOVES ? go.n : MAX_MOVES;
        if (lane < LINE_WORDS) A.node_line[(nbase + leaf) * LINE_WORDS + lane] = w;
        NodeMeta nm;
        nm.N = 1;
        nm.first_edge = -1;
        nm.ne_term = NODE_TERM;
        nm.val = 0.0f;
        const int n_edges = h->n_edges;
        if (n == 0) {
            nm.val = (go.flags & RF_CHECKMATE) ? -1.0f : 0.0f;
        } else if (go.flags & RF_ONLY_KINGS) {
            nm.val = 0.0f;
        } else if (n_edges + n > cfg.edge_cap) {
            nm.val = 0.0f;
            if (lane == 0) h->overflow |= HDR_OVERFLOW;
        } else {
            nm.first_edge = n_edges;
            nm.ne_term = n | NODE_PENDING;
            for (int k = lane; k < n; k += 32) {
                const size_t e = ebase + n_edges + k;
                A.eMv[e] = mv[k];
                A.eN[e] = 0;
                A.eW[e] = 0.0f;
                A.eP[e] = 0.0f;
                A.eChild[e] = -1;
            }
            int kind = 0, aux = -1;
            uint64_t ch = 0;
            if (cfg.cache_mask) {
                ch = cache_hash(pos_hash_warp(w, lane));
                kind = cache_lookup_warp(lane, cfg, A, wave, gs, w, ch, aux);
            }
            if (kind) {   // cache hit / in-wave follower: no tower pass, expanded by the late kernel
                uint32_t li = 0;
                if (lane == 0) li = atomic_add_u32(A.n_late, 1u);
                li = (uint32_t)shfl32((int)li, 0);
                if (lane == 0) {
                    A.late_game[li] = gs;
                    A.late_src[li] = aux;
                }
            } else {
                uint32_t slot = 0;
                if (lane == 0) slot = atomic_add_u32(A.n_eval, 1u);
                slot = (uint32_t)shfl32((int)slot, 0);
                if (lane < LINE_WORDS) A.eval_lines[(size_t)slot * LINE_WORDS + lane] = w;
                int ci = -1;
                if (cfg.cache_mask) ci = cache_claim_warp(lane, cfg, A, wave, w, ch, (int)slot);
                if (lane == 0) {
                    A.eval_game[slot] = gs;
                    if (cfg.cache_mask) {
                        A.eval_centry[slot] = ci;
                        A.eval_hash[slot] = ch;
                    }
                }
            }
            if (lane == 0) {
                h->n_edges = n_edges + n;
                A.pend_node[gs] = leaf;
                A.pend_depth[gs] = depth;
            }
            queued = true;
        }
        v = nm.val;
        if (lane == 0) {
            A.node_meta[nbase + leaf] = nm;
            h->n_nodes = n_nodes + 1;
        }
    }
    syncwarp();
    if (!queued) {
        mcts_backup_warp(lane, A, ebase, nbase, pbase, depth, v);
        if (lane == 0) h->sims_done += 1;
    }
    syncwarp();
    return queued ? 1 : 0;
}

// One search wave for game g: up to cfg.inflight selections, one after the other (each sees the virtual visits of
// the earlier ones), as long as the move's simulation budget allows.  A selection that runs into a leaf still
// waiting for its evaluation ends the wave for this game.
KV_DEV void mcts_select_warp(const Tables& T, int lane, const MctsCfg& cfg, const MctsArrays& A, int g, uint16_t* mv,
                             uint32_t wave = 0) {
    GameHdr* h = &A.hdr[g];
    if (h->done) return;
    int n_pend = 0;
    for (int j = 0; j < cfg.inflight; j++) {
        if (h->sims_done + n_pend >= cfg.sims) break;
        const int r = mcts_select_one_warp(T, lane, cfg, A, g, g * cfg.inflight + n_pend, mv, wave);
        if (r == 2) break;
        n_pend += r;
    }
    if (lane == 0) h->n_pend = n_pend;
    syncwarp();
}

// After the wave's expansions (K > 1): back the pending simulations of game g up in slot order.
KV_DEV void mcts_backup_game_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int g) {
    GameHdr* h = &A.hdr[g];
    const int np = h->n_pend;
    for (int j = 0; j < np; j++) mcts_commit_warp(lane, cfg, A, g, g * cfg.inflight + j);   // slot order: paths share edges
    if (lane == 0) h->n_pend = 0;
    syncwarp();
}

// Finish the pending simulation of game g given the leaf's legal-move logits (n floats, edge order) and the
// evaluator's white-perspective value: softmax priors (+ root Dirichlet noise), then backup.
// `logits` is per-warp scratch (shared memory on the device) and is overwritten.
// Softmax statistics over all POLICY_N logits L(i) (scripts/self_play.py:150 takes the softmax over every index):
// lane-strided maximum, then lane-strided sum of exp(l - max), both combined by an xor butterfly — the oracle walks
// the same order.  All lanes return the same (mx, z).
template <class LogitFn>
KV_DEV void full_softmax_stats_warp(int lane, LogitFn L, float& mx, float& z) {
    float m = -3.0e38f;
    for (int i = lane; i < POLICY_N; i += 32) {
        const float l = L(i);
        m = l > m ? l : m;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o = shfl_xorf(m, d, lane);
        m = o > m ? o : m;
    }
    float s = 0.0f;
    for (int i = lane; i < POLICY_N; i += 32) s = s + kvd_expf(L(i) - m);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s = s + shfl_xorf(s, d, lane);
    mx = m;
    z = s;
}

// Root priors by the reference's rule (cfg.root_mix; scripts/self_play.py:150-167): policy = softmax over all 4096
// logits, noise = Dirichlet(alpha) over all 4096 indices (Gamma variates keyed by seed, game, ply, INDEX), mixed with
// eps, then the legal entries renormalised.  logits[k] holds the legal moves' logits (edge order) and is overwritten.
// gsum_pre > 0: the sum of the 4096 Gamma variates, already formed by the caller (the network evaluator's CTA computes
// it with all its threads); otherwise the warp forms it here (lane-strided, butterfly: the order the oracle walks).
KV_DEV void mcts_root_mix_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, const GameHdr* h, size_t e0, int n,
                               float* logits, float mx_all, float z_all, float gsum_pre = 0.0f) {
    for (int k = lane; k < n; k += 32) logits[k] = kvd_expf(logits[k] - mx_all) / z_all;
    if (cfg.dir_eps > 0.0f) {
        const uint64_t key = (uint64_t)h->ply * POLICY_N;
        float part = gsum_pre;
        if (!(gsum_pre > 0.0f)) {
            part = 0.0f;
            for (int i = lane; i < POLICY_N; i += 32) part = part + kvd_gamma_small(cfg.dir_alpha, cfg.seed, h->game_id, key + (uint64_t)i);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) part = part + shfl_xorf(part, d, lane);
        }
        for (int k = lane; k < n; k += 32) {
            const float g = kvd_gamma_small(cfg.dir_alpha, cfg.seed, h->game_id, key + (uint64_t)move_index(A.eMv[e0 + k]));
            logits[k] = (1.0f - cfg.dir_eps) * logits[k] + cfg.dir_eps * (g / part);
        }
    }
    syncwarp();
    float s = 0.0f;
    if (lane == 0)
        for (int k = 0; k < n; k++) s = s + logits[k];
    s = shflf(s, 0);
    for (int k = lane; k < n; k += 32) A.eP[e0 + k] = logits[k] / s;
}

KV_DEV void mcts_expand_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int gs, float* logits, float v_white,
                             bool from_cache = false, float mx_all = 0.0f, float z_all = 0.0f, float gsum_pre = 0.0f) {
    const int g = gs / cfg.inflight;
    GameHdr* h = &A.hdr[g];
    const size_t nbase = (size_t)g * cfg.node_cap, ebase = (size_t)g * cfg.edge_cap;
    const int c = A.pend_node[gs];
    const NodeMeta m = A.node_meta[nbase + c];
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = ebase + m.first_edge;
    const bool wtm = A.node_line[(nbase + c) * LINE_WORDS + 12] & 1;
    if (c == 0 && cfg.root_mix) {
        mcts_root_mix_warp(lane, cfg, A, h, e0, n, logits, mx_all, z_all, gsum_pre);
    } else {
    float mx = -3.0e38f;
    for (int k = lane; k < n; k += 32) mx = logits[k] > mx ? logits[k] : mx;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const float o = shfl_xorf(mx, d, lane);
        mx = o > mx ? o : mx;
    }
    for (int k = lane; k < n; k += 32) logits[k] = kvd_expf(logits[k] - mx);
    syncwarp();
    float s = 0.0f;
    if (lane == 0)
        for (int k = 0; k < n; k++) s = s + logits[k];   // sequential order = the oracle's
    s = shflf(s, 0);
    for (int k = lane; k < n; k += 32) A.eP[e0 + k] = logits[k] / s;
    if (c == 0 && cfg.dir_eps > 0.0f) {
        // root Dirichlet(alpha) noise over the legal moves (scripts/self_play.py:153-154 mixes with eps = 0.25)
        for (int k = lane; k < n; k += 32)
            A.eW[e0 + k] = kvd_gamma_small(cfg.dir_alpha, cfg.seed, h->game_id, (uint64_t)h->ply * 256 + (uint64_t)k);
        syncwarp();
        float gsum = 0.0f;
        if (lane == 0)
            for (int k = 0; k < n; k++) gsum = gsum + A.eW[e0 + k];
        gsum = shflf(gsum, 0);
        for (int k = lane; k < n; k += 32) {
            const float eta = A.eW[e0 + k] / gsum;
            A.eP[e0 + k] = (1.0f - cfg.dir_eps) * A.eP[e0 + k] + cfg.dir_eps * eta;
            A.eW[e0 + k] = 0.0f;
        }
    }
    }
    syncwarp();
    const float v = wtm ? v_white : -v_white;
    if (lane == 0) {
        A.node_meta[nbase + c].val = v;
        A.node_meta[nbase + c].ne_term = n;   // priors delivered: no longer pending
        A.pend_kind[gs] = from_cache ? 2 : 1;
    }
    syncwarp();
    // K == 1: the game's only in-flight simulation, back it up here; K > 1: mcts_backup_game_warp does it in
    // slot order once every expansion of the wave is in (other CTAs may be expanding this game's other leaves)
    if (cfg.inflight == 1) {
        mcts_commit_warp(lane, cfg, A, g, gs);
        if (lane == 0) h->n_pend = 0;
        syncwarp();
    }
}

// Hash evaluator (test evaluator, oracle mode 0): logits and value from a hash of the position.
// scratch: n floats of per-warp scratch (shared memory on the device).
KV_DEV void mcts_hash_eval_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int slot, float* scratch,
                                uint32_t wave = 0) {
    const int gs = A.eval_game[slot], g = gs / cfg.inflight;
    const uint64_t w = lane < LINE_WORDS ? A.eval_lines[(size_t)slot * LINE_WORDS + lane] : 0ull;
    const uint64_t ph = pos_hash_warp(w, lane);
    if (cfg.cache_mask) cache_fill_warp(lane, cfg, A, wave, slot, nullptr, hash_value(ph));
    const NodeMeta m = A.node_meta[(size_t)g * cfg.node_cap + A.pend_node[gs]];
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = (size_t)g * cfg.edge_cap + m.first_edge;
    for (int k = lane; k < n; k += 32) scratch[k] = hash_logit(ph, move_index(A.eMv[e0 + k]));
    syncwarp();
    float mx_all = 0.0f, z_all = 0.0f;
    if (cfg.root_mix && A.pend_node[gs] == 0)
        full_softmax_stats_warp(lane, [&](int i) { return hash_logit(ph, i); }, mx_all, z_all);
    mcts_expand_warp(lane, cfg, A, gs, scratch, hash_value(ph), false, mx_all, z_all);
}

// Hash-evaluator counterpart of the late (cache hit / follower) expansion: the value comes from the cached / leader
// features, the logits are recomputed from the position hash (what a miss would have produced, bit for bit).
KV_DEV void mcts_hash_late_warp(int lane, const MctsCfg& cfg, const MctsArrays& A, int li, float* scratch) {
    const int gs = A.late_game[li], src = A.late_src[li], g = gs / cfg.inflight;
    const size_t nidx = (size_t)g * cfg.node_cap + A.pend_node[gs];
    const uint64_t w = lane < LINE_WORDS ? A.node_line[nidx * LINE_WORDS + lane] : 0ull;
    const uint64_t ph = pos_hash_warp(w, lane);
    const NodeMeta m = A.node_meta[nidx];
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = (size_t)g * cfg.edge_cap + m.first_edge;
    for (int k = lane; k < n; k += 32) scratch[k] = hash_logit(ph, move_index(A.eMv[e0 + k]));
    syncwarp();
    const float v = src < 0 ? A.feat_game[(size_t)gs * FEAT + 128] : A.feat_slot[(size_t)src * FEAT + 128];
    float mx_all = 0.0f, z_all = 0.0f;
    if (cfg.root_mix && A.pend_node[gs] == 0)
        full_softmax_stats_warp(lane, [&](int i) { return hash_logit(ph, i); }, mx_all, z_all);
    mcts_expand_warp(lane, cfg, A, gs, scratch, v, true, mx_all, z_all);
}

// After cfg.sims simulations: pick the move from the root visit counts, record (position, move), play it,
// detect the end of the game (scripts/self_play.py:122-238 control flow) and reset the tree.
KV_DEV void mcts_finish_move_warp(const Tables& T, int lane, const MctsCfg& cfg, const MctsArrays& A, int g, uint16_t* mv) {
    GameHdr* h = &A.hdr[g];
    if (h->done) return;
    const size_t nbase = (size_t)g * cfg.node_cap, ebase = (size_t)g * cfg.edge_cap;
    const NodeMeta m = A.node_meta[nbase];
    const int ply = h->ply;
    uint64_t w = lane < LINE_WORDS ? A.root_line[(size_t)g * LINE_WORDS + lane] : 0ull;
    if (h->n_nodes == 0 || (m.ne_term & NODE_TERM)) {   // root had no move at all: the game is over as it stands
        if (lane == 0) {
            h->done = 1;
            h->result = (h->n_nodes && m.val < 0.0f) ? ((w & 1) ? -1 : 1) : 0;
        }
        syncwarp();
        return;
    }
    const int n = m.ne_term & 0xFFFF;
    const size_t e0 = ebase + m.first_edge;
    // the value the resign rule looks at: the evaluator's output for the position the move is chosen in, as the
    // network returned it (m.val is from the side to move's perspective; the flip is exact) — or the scripted one
    const bool wtm_root = shfl64(w, 12) & 1;
    float v_resign = wtm_root ? m.val : -m.val;
    int forced = -1;
    if (cfg.script_stride > 0 && ply < cfg.script_stride) {
        const size_t si = (size_t)g * cfg.script_stride + ply;
        if (A.script_val) {
            const float sv = A.script_val[si];
            if (sv == sv) v_resign = sv;
        }
        const int want = A.script_move ? (int)A.script_move[si] : 0xFFFF;
        if (want != 0xFFFF) {
            uint32_t hit = 0;
            for (int k0 = 0; k0 < n; k0 += 32) {   // warp-uniform trip count
                const int k = k0 + lane;
                hit = ballot(k < n && ((int)A.eMv[e0 + k] & 0xFFF) == (want & 0xFFF));
                if (hit) {
                    forced = k0 + ffs32(hit) - 1;
                    break;
                }
            }
            if (forced < 0) {   // not a legal move here: stop the game (draw) and flag it
                if (lane == 0) {
                    h->done = 1;
                    h->result = 0;
                    h->overflow |= HDR_SCRIPT_MISS;
                }
                syncwarp();
                return;
            }
        }
    }
    // total visits
    int tot = 0;
    for (int k = lane; k < n; k += 32) tot += (int)A.eN[e0 + k];
    tot = warp_sum32(tot, lane);
    int pick = 0;
    if (forced >= 0) {
        pick = forced;
    } else if (tot == 0) {
        // sims == 1: only the root was expanded.  Sample the move from the (noisy) priors — with cfg.root_mix the
        // distribution of the reference's own move rule (scripts/self_play.py:147-167: softmax and Dirichlet noise
        // over all 4096 indices, legal renormalise, sample; the generator differs from numpy / random.choices).
        // Sequential fp32 prefix sum by lane 0, same order as the oracle.
        if (lane == 0) {
            const float u = kvd_u01(kvd_rand24(cfg.seed, h->game_id, (uint64_t)ply, 0xC0FFEEull));
            float sum = 0.0f;
            for (int k = 0; k < n; k++) sum = sum + A.eP[e0 + k];
            const float thr = u * sum;
            float cum = 0.0f;
            pick = n - 1;
            for (int k = 0; k < n; k++) {
                cum = cum + A.eP[e0 + k];
                if (cum > thr) {
                    pick = k;
                    break;
                }
            }
        }
        pick = shfl32(pick, 0);
    } else if (ply < cfg.temp_plies) {
        const uint64_t r = ((uint64_t)kvd_rand24(cfg.seed, h->game_id, (uint64_t)ply, 0xC0FFEEull) * (uint64_t)tot) >> 24;
        int base = 0;
        pick = -1;
        for (int k0 = 0; k0 < n; k0 += 32) {   // warp-uniform trip count
            const int k = k0 + lane;
            const int c = k < n ? (int)A.eN[e0 + k] : 0;
            const int incl = warp_incl_scan(c, lane) + base;
            const uint32_t hit = ballot(k < n && (uint64_t)incl > r);
            if (pick < 0 && hit) pick = k0 + ffs32(hit) - 1;
            base = shfl32(incl, 31);
        }
        if (pick < 0) pick = 0;
    } else {
        int bc = -1, bi = 0x7FFFFFFF;
        for (int k = lane; k < n; k += 32) {
            const int c = (int)A.eN[e0 + k];
            if (c > bc) {
                bc = c;
                bi = k;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const int oc = shfl_xor32(bc, d, lane), oi = shfl_xor32(bi, d, lane);
            if (oc > bc || (oc == bc && oi < bi)) {
                bc = oc;
                bi = oi;
            }
        }
        pick = bi;
    }
    const int mvw = A.eMv[e0 + pick];
    if (ply < cfg.rec_cap) {
        if (lane < 12) A.rec_line[((size_t)g * cfg.rec_cap + ply) * 12 + lane] = w;
        if (lane == 0) A.rec_move[(size_t)g * cfg.rec_cap + ply] = (uint16_t)mvw;
    }
    w = make_move_warp(lane, w, mvw, T_Q);
    // isDraw() (:21-33, only kings left) is asked right after makeMove, before the next getValidMoves can rewrite the board
    const bool only_kings = ballot(lane < 12 && lane != 0 && lane != 6 && w != 0ull) == 0;
    const GenOut go = movegen_warp(T, lane, w, mv);
    if (lane < LINE_WORDS) A.root_line[(size_t)g * LINE_WORDS + lane] = w;
    const bool wtm_new = shfl64(w, 12) & 1;
    if (lane == 0) {
        const int np = ply + 1;
        h->ply = np;
        h->n_nodes = 0;
        h->n_edges = 0;
        h->sims_done = 0;
        h->n_pend = 0;
        // the reference's tests after a move, in its order (scripts/self_play.py:180-199), then the loop head (:125)
        if (only_kings) {                                     // isDraw() :180 -> :225
            h->done = 1;
            h->result = 0;
        } else if (cfg.resign_min_plies >= 0 && np > cfg.resign_min_plies && v_resign < cfg.resign_thr) {
            h->done = 1;                                      // resignation :185-189
            h->result = wtm_new ? -1 : 1;
        } else if (np >= cfg.max_plies) {                     // max_moves :196-199 -> :209-211 (even on a mating move)
            h->done = 1;
            h->result = 0;
        } else if (go.n == 0) {                               // :125 -> checkmate :217-220 / stalemate :221-224
            h->done = 1;
            h->result = (go.flags & RF_CHECKMATE) ? (wtm_new ? -1 : 1) : 0;
        }
    }
    syncwarp();
}

}  // namespace kv
