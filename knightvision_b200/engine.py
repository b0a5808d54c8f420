"""Batched device engine: thin Python host over the C ABI (include/kv_b200.h).

torch is used for device memory and streams only; every computation is a hand-written sm_100a kernel
inside libkv_b200.so.  Board lines are uint64 [n,16] (see knightvision_b200.layout).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as N

MOVE_STRIDE = 256


def _ptr(t):
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return t.ctypes.data_as(ctypes.c_void_p)


def lines_to_device(lines: np.ndarray, device) -> torch.Tensor:
    a = np.ascontiguousarray(lines, dtype=np.uint64).view(np.int64)
    return torch.from_numpy(a).to(device)


def lines_to_host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


class Engine:
    """One context per GPU (kv_create).  Raises if there is no CUDA device or the library is missing."""

    def __init__(self, device: int | str | torch.device = 0):
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise N.KVError("knightvision_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = dev
        self.index = dev.index or 0
        self._lib = N.lib()
        ctx = ctypes.c_void_p()
        rc = self._lib.kv_create(self.index, ctypes.byref(ctx))
        if rc != 0:
            raise N.KVError(f"kv_create failed ({rc}): {self._lib.kv_last_error(None).decode()}")
        self.ctx = ctx
        torch.cuda.set_device(dev)

    def close(self):
        if getattr(self, "ctx", None):
            self._lib.kv_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        return int(self._lib.kv_launch_count(self.ctx))

    @property
    def sm_count(self) -> int:
        return int(self._lib.kv_sm_count(self.ctx))

    KERNELS = ("movegen", "make_moves", "perft_expand", "perft_leaf", "encode", "net_stem", "net_conv", "net_head",
               "mcts_select", "mcts_expand", "mcts_misc", "train_wgrad", "train_bn")

    def profile(self, on: bool):
        N.check(self.ctx, self._lib.kv_profile_enable(self.ctx, int(on)), "kv_profile_enable")

    def profile_read(self) -> dict:
        """{kernel: (total_ms, launches)} recorded with CUDA events since the last read."""
        ms = np.zeros(16, dtype=np.float64)
        n = np.zeros(16, dtype=np.uint64)
        k = self._lib.kv_profile_read(self.ctx, _ptr(ms), _ptr(n), 16)
        return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(self.KERNELS[:k])}

    # ---- rules ---------------------------------------------------------------------------------
    def movegen(self, lines: torch.Tensor, moves=None, counts=None, flags=None, stride: int = MOVE_STRIDE):
        """getValidMoves for every board.  lines: int64/uint64 view [n,16] on this device (rewritten in
        place only in the RF_STATE_MUTATED corner).  Returns (moves int16 [n,stride], counts, flags)."""
        n = lines.shape[0]
        if moves is None:
            moves = torch.empty((n, stride), dtype=torch.int16, device=self.device)
        if counts is None:
            counts = torch.empty(n, dtype=torch.int32, device=self.device)
        if flags is None:
            flags = torch.empty(n, dtype=torch.int32, device=self.device)
        N.check(self.ctx, self._lib.kv_movegen(self.ctx, _ptr(lines), n, _ptr(moves), stride, _ptr(counts),
                                               _ptr(flags), self._stream()), "kv_movegen")
        return moves, counts, flags

    def make_moves(self, lines: torch.Tensor, mv: torch.Tensor):
        """makeMove in place; mv int16 [n] move words (0xFFFF = skip)."""
        N.check(self.ctx, self._lib.kv_make_moves(self.ctx, _ptr(lines), lines.shape[0], _ptr(mv), self._stream()),
                "kv_make_moves")
        return lines

    def perft(self, roots: torch.Tensor, depth: int, chunk: int = 0) -> torch.Tensor:
        n = roots.shape[0]
        out = torch.empty((n, 8), dtype=torch.int64, device=self.device)
        N.check(self.ctx, self._lib.kv_perft(self.ctx, _ptr(roots), n, depth, _ptr(out), chunk, self._stream()),
                "kv_perft")
        return out

    def encode(self, lines: torch.Tensor) -> torch.Tensor:
        n = lines.shape[0]
        out = torch.empty((n, 12, 8, 8), dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_encode(self.ctx, _ptr(lines), n, _ptr(out), self._stream()), "kv_encode")
        return out

    # ---- network ---------------------------------------------------------------------------------
    def net_create(self, stem: int = 256, tower: int = 512, blocks: int = 5, conv2: bool = True, max_boards: int = 4096):
        N.check(self.ctx, self._lib.kv_net_create(self.ctx, stem, tower, blocks, int(conv2), max_boards), "kv_net_create")
        self.net_max_boards = max_boards
        self.net_geometry = None          # bookkeeping of ChessNet.attach (re-created only when the shape changes)

    def net_set_tower_fused(self, mode):
        """0 = one launch per tower layer, 1 / True = whole tower as one dependency-scheduled launch (default, same bits),
        2 = whole-tower launch with the halo activation operand (default; 3 instead of 9 fetches per tile; fp32-rounding
        equal), 3 / 4 = mode 2 on 4-CTA clusters sharing the weight tiles by multicast (4: plus CTA pairs on the SMs the
        clusters cannot cover; same bits as 2)."""
        N.check(self.ctx, self._lib.kv_net_set_tower_fused(self.ctx, int(mode)), "kv_net_set_tower_fused")

    def net_tower_clusters4(self) -> int:
        """Co-resident 4-CTA clusters of the weight-multicast tower kernel (0: net_set_tower_fused(3) unavailable)."""
        return int(self._lib.kv_net_tower_clusters4(self.ctx))

    def net_set_conv_mode(self, cta_group: int):
        N.check(self.ctx, self._lib.kv_net_set_conv_mode(self.ctx, cta_group), "kv_net_set_conv_mode")

    def net_load(self, blob: torch.Tensor):
        """blob: fp32 CPU tensor, ChessNet.weight_blob()."""
        blob = blob.detach().to(torch.float32).cpu().contiguous()
        need = int(self._lib.kv_net_blob_floats(self.ctx))
        if blob.numel() != need:
            raise N.KVError(f"weight blob has {blob.numel()} floats, the architecture needs {need}")
        N.check(self.ctx, self._lib.kv_net_load(self.ctx, ctypes.c_void_p(blob.data_ptr()), blob.numel()), "kv_net_load")

    def net_blob_tensor(self) -> torch.Tensor:
        """Device staging buffer for the fp32 weight blob as a torch tensor (NCCL broadcast target)."""
        n = int(self._lib.kv_net_blob_floats(self.ctx))
        ptr = int(self._lib.kv_net_blob_device_ptr(self.ctx))

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Holder(), device=self.device)

    def net_folded_tensor(self) -> torch.Tensor:
        """The folded weights the kernels read (bf16 tower, fp32 biases / stem table / heads) as one uint8 device tensor:
        the NCCL broadcast payload between generations (half the fp32 blob); receivers call net_adopt_folded()."""
        n = int(self._lib.kv_net_folded_bytes(self.ctx))
        ptr = int(self._lib.kv_net_folded_device_ptr(self.ctx))

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Holder(), device=self.device)

    def net_adopt_folded(self):
        N.check(self.ctx, self._lib.kv_net_adopt_folded(self.ctx, self._stream()), "kv_net_adopt_folded")

    def net_commit(self):
        N.check(self.ctx, self._lib.kv_net_commit_weights(self.ctx, self._stream()), "kv_net_commit_weights")

    def net_forward(self, lines: torch.Tensor, want_policy: bool = True):
        n = lines.shape[0]
        pol = torch.empty((n, 4096), dtype=torch.float32, device=self.device) if want_policy else None
        val = torch.empty(n, dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_net_forward(self.ctx, _ptr(lines), n, _ptr(pol) if want_policy else None,
                                                   _ptr(val), self._stream()), "kv_net_forward")
        return pol, val

    def net_forward_partial(self, lines: torch.Tensor, n_convs: int) -> torch.Tensor:
        """Test hook: NHWC bf16 activations [n,8,8,C] after the stem and the first n_convs tower convolutions."""
        n = lines.shape[0]
        out = torch.empty((n, 8, 8, 512), dtype=torch.bfloat16, device=self.device)
        ch = ctypes.c_int(0)
        N.check(self.ctx, self._lib.kv_net_forward_partial(self.ctx, _ptr(lines), n, n_convs, _ptr(out), ctypes.byref(ch)),
                "kv_net_forward_partial")
        return out.reshape(-1)[: n * 64 * ch.value].reshape(n, 8, 8, ch.value)

    def net_forward_planes(self, planes: torch.Tensor):
        n = planes.shape[0]
        pol = torch.empty((n, 4096), dtype=torch.float32, device=self.device)
        val = torch.empty(n, dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_net_forward_planes(self.ctx, _ptr(planes), n, _ptr(pol), _ptr(val),
                                                          self._stream()), "kv_net_forward_planes")
        return pol, val

    # ---- training-side convolution operators (kv_train.cu) ------------------------------------------------------
    def conv3x3_pack(self, weight: torch.Tensor, flip_transpose: bool = False) -> torch.Tensor:
        """fp32 [Cout,Cin,3,3] parameter -> bf16 [Cout,9,Cin] (or the dgrad operand [Cin,9,Cout], taps mirrored)."""
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        w = weight.detach().to(torch.float32).contiguous()
        out = torch.empty((cin, 9, cout) if flip_transpose else (cout, 9, cin), dtype=torch.bfloat16, device=self.device)
        N.check(self.ctx, self._lib.kv_conv3x3_pack(self.ctx, _ptr(w), cout, cin, int(flip_transpose), _ptr(out),
                                                    self._stream()), "kv_conv3x3_pack")
        return out

    def conv3x3_fprop(self, x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor | None = None,
                      residual: torch.Tensor | None = None, relu: bool = False) -> torch.Tensor:
        """x: bf16 [n,8,8,Cin] contiguous (NHWC); w_packed [Cout,9,Cin] bf16.  Returns bf16 [n,8,8,Cout]."""
        n, cin, cout = int(x.shape[0]), int(x.shape[3]), int(w_packed.shape[0])
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and tuple(x.shape[1:3]) == (8, 8)
        assert w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous() and int(w_packed.shape[2]) == cin
        y = torch.empty((n, 8, 8, cout), dtype=torch.bfloat16, device=self.device)
        N.check(self.ctx, self._lib.kv_conv3x3_fprop(self.ctx, _ptr(x), _ptr(w_packed),
                                                     _ptr(bias) if bias is not None else None,
                                                     _ptr(residual) if residual is not None else None, _ptr(y), n, cin, cout,
                                                     int(relu), self._stream()), "kv_conv3x3_fprop")
        return y

    def conv3x3_wgrad(self, x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
        """x bf16 [n,8,8,Cin], dy bf16 [n,8,8,Cout] (NHWC contiguous) -> fp32 [Cout,Cin,3,3]."""
        n, cin, cout = int(x.shape[0]), int(x.shape[3]), int(dy.shape[3])
        assert x.dtype == dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_conv3x3_wgrad(self.ctx, _ptr(x), _ptr(dy), _ptr(dw), n, cin, cout, self._stream()),
                "kv_conv3x3_wgrad")
        return dw

    def bn_relu_fwd(self, z: torch.Tensor, gamma, beta, running_mean, running_var, momentum: float, eps: float,
                    residual: torch.Tensor | None = None, relu: bool = True):
        """Train-mode BatchNorm + ReLU (+ residual) on NHWC bf16 [n,8,8,C]; returns (y, save_mean, save_rstd)."""
        C = int(z.shape[-1]); rows = z.numel() // C
        assert z.dtype == torch.bfloat16 and z.is_contiguous()
        y = torch.empty_like(z)
        mean = torch.empty(C, dtype=torch.float32, device=self.device)
        rstd = torch.empty(C, dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_bn_relu_fwd(
            self.ctx, _ptr(z), _ptr(residual) if residual is not None else None, _ptr(gamma), _ptr(beta),
            _ptr(running_mean) if running_mean is not None else None, _ptr(running_var) if running_var is not None else None,
            momentum, eps, _ptr(y), _ptr(mean), _ptr(rstd), rows, C, int(relu), self._stream()), "kv_bn_relu_fwd")
        return y, mean, rstd

    def bn_relu_bwd(self, dy, y, z, gamma, mean, rstd, relu: bool = True, want_dres: bool = False):
        """Returns (dz bf16, dres bf16 or None, dgamma fp32, dbeta fp32)."""
        C = int(z.shape[-1]); rows = z.numel() // C
        assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and y.is_contiguous() and z.is_contiguous()
        dz = torch.empty_like(z)
        dres = torch.empty_like(z) if want_dres else None
        dg = torch.empty(C, dtype=torch.float32, device=self.device)
        db = torch.empty(C, dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_bn_relu_bwd(self.ctx, _ptr(dy), _ptr(y), _ptr(z), _ptr(gamma), _ptr(mean), _ptr(rstd),
                                                   _ptr(dz), _ptr(dres) if dres is not None else None, _ptr(dg), _ptr(db),
                                                   rows, C, int(relu), self._stream()), "kv_bn_relu_bwd")
        return dz, dres, dg, db

    def channel_sum(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 [C] column sums of an NHWC bf16 tensor (the convolution-bias gradient)."""
        C = int(x.shape[-1]); rows = x.numel() // C
        assert x.dtype == torch.bfloat16 and x.is_contiguous()
        out = torch.empty(C, dtype=torch.float32, device=self.device)
        N.check(self.ctx, self._lib.kv_channel_sum(self.ctx, _ptr(x), _ptr(out), rows, C, self._stream()), "kv_channel_sum")
        return out

    # ---- self-play search ---------------------------------------------------------------------------
    def mcts_create(self, n_games: int, sims: int, max_plies: int, temp_plies: int = 30, c_puct: float = 1.5,
                    dir_alpha: float = 0.3, dir_eps: float = 0.25, seed: int = 42, eval_mode: int = 1,
                    edges_per_node: int = 0, inflight: int = 1):
        """inflight = K simulations in flight per game and wave (virtual loss when K > 1); the network must have been
        created with max_boards >= n_games * K."""
        N.check(self.ctx, self._lib.kv_mcts_create_k(self.ctx, n_games, sims, edges_per_node, max_plies, temp_plies,
                                                     c_puct, dir_alpha, dir_eps, seed, eval_mode, inflight),
                "kv_mcts_create_k")
        self.mcts_inflight = inflight
        self.mcts_geometry_key = None     # bookkeeping of SelfPlay (pools are kept while the geometry is unchanged)
        self.mcts_cache_log2 = 0
        g = np.zeros(4, dtype=np.int32)
        N.check(self.ctx, self._lib.kv_mcts_geometry(self.ctx, _ptr(g)), "kv_mcts_geometry")
        self.mcts_games, self.mcts_node_cap, self.mcts_edge_cap, self.mcts_rec_cap = (int(x) for x in g)
        self.mcts_sims = sims

    def mcts_enable_cache(self, log2_slots: int):
        """Evaluation cache of 2**log2_slots x 640 B entries (0 = off); results are identical with it on or off."""
        N.check(self.ctx, self._lib.kv_mcts_enable_cache(self.ctx, log2_slots), "kv_mcts_enable_cache")
        self.mcts_cache_log2 = log2_slots

    def mcts_reset(self, start_lines: torch.Tensor | None = None, game_id_base: int = 0):
        N.check(self.ctx, self._lib.kv_mcts_reset(self.ctx, _ptr(start_lines) if start_lines is not None else None,
                                                  game_id_base, self._stream()), "kv_mcts_reset")

    def mcts_run_sims(self, n_waves: int):
        N.check(self.ctx, self._lib.kv_mcts_run_sims(self.ctx, n_waves, self._stream()), "kv_mcts_run_sims")

    def mcts_finish_move(self):
        N.check(self.ctx, self._lib.kv_mcts_finish_move(self.ctx, self._stream()), "kv_mcts_finish_move")

    def mcts_run_move(self):
        N.check(self.ctx, self._lib.kv_mcts_run_move(self.ctx, self._stream()), "kv_mcts_run_move")

    def mcts_set_pipeline(self, mode: int = -1):
        """Two game groups whose waves alternate on two streams (tree kernels of one group under the other group's
        tower): -1 default (off unless KV_MCTS_PIPELINE=1), 0 off, 1 on.  Search results are identical either way."""
        N.check(self.ctx, self._lib.kv_mcts_set_pipeline(self.ctx, mode), "kv_mcts_set_pipeline")

    def mcts_set_eval_split(self, mode: int = -1):
        """Evaluator schedule after the tower: 0 one kernel per leaf (default), 1 head-features kernel + batched finish
        kernel.  Results are identical either way (and so is the power-capped step time)."""
        N.check(self.ctx, self._lib.kv_mcts_set_eval_split(self.ctx, mode), "kv_mcts_set_eval_split")

    def mcts_waves(self) -> int:
        """Search waves launched since mcts_create (with K > 1 a move takes a data-dependent number of waves)."""
        return int(self._lib.kv_mcts_waves(self.ctx))

    def mcts_set_resign(self, threshold: float = -0.7, min_plies: int = 15):
        """Resignation rule of scripts/self_play.py:184-189 (defaults = the reference's); min_plies < 0 = off."""
        N.check(self.ctx, self._lib.kv_mcts_set_resign(self.ctx, threshold, min_plies), "kv_mcts_set_resign")

    def mcts_set_root_mix(self, mode: int = -1):
        """Root priors: 0 legal-only softmax + noise, 1 the reference's mixing over all 4096 indices
        (scripts/self_play.py:150-167), -1 default (1 when sims == 1)."""
        N.check(self.ctx, self._lib.kv_mcts_set_root_mix(self.ctx, mode), "kv_mcts_set_root_mix")

    def mcts_set_script(self, moves: torch.Tensor | None = None, values: torch.Tensor | None = None):
        """Scripted play: moves int16 [n_games, stride] move words (-1 = 0xFFFF = choose as usual), values float32
        [n_games, stride] (NaN = the evaluator's value) on this device; both None clears the script."""
        stride = 0
        for t in (moves, values):
            if t is not None:
                assert t.is_contiguous() and t.shape[0] == self.mcts_games and t.device == self.device
                stride = int(t.shape[1])
        self._script = (moves, values)     # keep the tensors alive while the context points at them
        N.check(self.ctx, self._lib.kv_mcts_set_script(self.ctx, _ptr(moves) if moves is not None else None,
                                                       _ptr(values) if values is not None else None, stride),
                "kv_mcts_set_script")

    def mcts_status(self) -> dict:
        out = np.zeros(10, dtype=np.uint64)
        N.check(self.ctx, self._lib.kv_mcts_status(self.ctx, _ptr(out), self._stream()), "kv_mcts_status")
        keys = ("done", "sims_in_move", "evals", "plies", "overflow", "white_wins", "black_wins", "draws", "cache_hits",
                "script_misses")
        return {k: int(v) for k, v in zip(keys, out)}

    def mcts_roots(self) -> torch.Tensor:
        """Current position of every game, int64 [n_games,16] on the device."""
        out = torch.empty((self.mcts_games, 16), dtype=torch.int64, device=self.device)
        N.check(self.ctx, self._lib.kv_mcts_get_roots(self.ctx, _ptr(out), self._stream()), "kv_mcts_get_roots")
        return out

    def mcts_cache_clear(self):
        N.check(self.ctx, self._lib.kv_mcts_cache_clear(self.ctx, self._stream()), "kv_mcts_cache_clear")

    def random_positions(self, n: int, max_plies: int = 40, seed: int = 1234) -> torch.Tensor:
        """n positions after k in [0, max_plies) uniformly random legal plies from the initial position (rules on the
        device; torch only draws the random numbers).  Positions without a legal move fall back to the initial one."""
        from . import layout as L
        g = torch.Generator(device="cpu").manual_seed(seed)
        start = lines_to_device(np.stack([L.start_line()] * n), self.device)
        lines = start.clone()
        k = torch.randint(0, max_plies, (n,), generator=g).to(self.device)
        for ply in range(max_plies):
            moves, counts, flags = self.movegen(lines)
            u = torch.rand(n, generator=g).to(self.device)
            idx = torch.clamp((u * counts.clamp(min=1)).long(), max=255)
            pick = moves.gather(1, idx[:, None])[:, 0]
            skip = (counts == 0) | (k <= ply) | ((flags & 16) != 0)
            pick = torch.where(skip, torch.full_like(pick, -1), pick)     # 0xFFFF = leave the board untouched
            self.make_moves(lines, pick.contiguous())
        moves, counts, flags = self.movegen(lines)
        bad = (counts == 0) | ((flags & 16) != 0)
        lines[bad] = start[bad]
        lines[:, 13:] = 0
        return lines

    def mcts_read_root(self, game: int) -> dict:
        mv = np.zeros(256, np.uint16); n_ = np.zeros(256, np.uint32); w = np.zeros(256, np.float32)
        p = np.zeros(256, np.float32); info = np.zeros(4, np.int32)
        N.check(self.ctx, self._lib.kv_mcts_read_root(self.ctx, game, _ptr(mv), _ptr(n_), _ptr(w), _ptr(p), _ptr(info)),
                "kv_mcts_read_root")
        n = int(info[0])
        return dict(moves=mv[:n], N=n_[:n], W=w[:n], P=p[:n], nodes=int(info[1]), edges=int(info[2]), ply=int(info[3]))

    def mcts_dump_tree(self, game: int):
        nv = np.zeros(self.mcts_node_cap, np.float32); nf = np.zeros(self.mcts_node_cap, np.int32)
        ep = np.zeros(self.mcts_edge_cap, np.float32); root = np.zeros(16, np.uint64)
        N.check(self.ctx, self._lib.kv_mcts_dump_tree(self.ctx, game, _ptr(nv), _ptr(nf), _ptr(ep), _ptr(root)),
                "kv_mcts_dump_tree")
        return nv, nf, ep, root

    def mcts_records(self):
        """(lines int64 [N,16], move_index int32 [N], reward float32 [N], game int32 [N]) device tensors, game order."""
        cnt = ctypes.c_int32(0)
        N.check(self.ctx, self._lib.kv_mcts_records(self.ctx, None, None, None, None, 0, ctypes.byref(cnt), self._stream()),
                "kv_mcts_records")
        n = int(cnt.value)
        lines = torch.zeros((max(n, 1), 16), dtype=torch.int64, device=self.device)
        move = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        reward = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device)
        game = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        if n:
            N.check(self.ctx, self._lib.kv_mcts_records(self.ctx, _ptr(lines), _ptr(move), _ptr(reward), _ptr(game), n,
                                                        ctypes.byref(cnt), self._stream()), "kv_mcts_records")
        return lines[:n], move[:n], reward[:n], game[:n]

    # ---- host-buffer forms (numpy in / numpy out; copies happen inside the C call) ---------------
    def movegen_host(self, lines: np.ndarray, stride: int = MOVE_STRIDE):
        lines = np.ascontiguousarray(lines, dtype=np.uint64).copy()
        n = lines.shape[0]
        moves = np.zeros((n, stride), dtype=np.uint16)
        counts = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        N.check(self.ctx, self._lib.kv_movegen_host(self.ctx, _ptr(lines), n, _ptr(moves), stride, _ptr(counts),
                                                    _ptr(flags)), "kv_movegen_host")
        return moves, counts, flags, lines

    def make_moves_host(self, lines: np.ndarray, mv: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(lines, dtype=np.uint64).copy()
        mv = np.ascontiguousarray(mv, dtype=np.uint16)
        N.check(self.ctx, self._lib.kv_make_moves_host(self.ctx, _ptr(out), out.shape[0], _ptr(mv)),
                "kv_make_moves_host")
        return out

    def attacked_host(self, lines: np.ndarray) -> np.ndarray:
        """squareUnderAttack over all 64 squares per board: uint64 masks (bit r*8+c)."""
        lines = np.ascontiguousarray(lines, dtype=np.uint64)
        out = np.zeros(lines.shape[0], dtype=np.uint64)
        N.check(self.ctx, self._lib.kv_attacked_host(self.ctx, _ptr(lines), lines.shape[0], _ptr(out)), "kv_attacked_host")
        return out

    def perft_host(self, roots: np.ndarray, depth: int, chunk: int = 0) -> np.ndarray:
        roots = np.ascontiguousarray(roots, dtype=np.uint64)
        out = np.zeros((roots.shape[0], 8), dtype=np.uint64)
        N.check(self.ctx, self._lib.kv_perft_host(self.ctx, _ptr(roots), roots.shape[0], depth, _ptr(out), chunk),
                "kv_perft_host")
        return out
