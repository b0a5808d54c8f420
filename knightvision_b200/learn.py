"""The caller of the hot path: the reference's RL loop (scripts/learn.py:152-209) and training step
(scripts/train.py:126-196), importable without tensorflow / google.colab / telegram / python-chess.

    train -> self-play -> dataset.extend -> train -> ...

Self-play runs on the B200 engine (knightvision_b200.selfplay); its records stay on the GPU in packed form
(12 bitboards + move + reward = 104 B per position instead of the reference's 3 080 B float planes) and are expanded
to planes batch by batch with the encode kernel.  The training step is PyTorch autograd over the
`ChessNet` parameter container, with the tower's 3x3 convolutions (99.8 % of the FLOPs) running forward, dgrad and
wgrad on the hand-written tcgen05 kernels (train_ops.py / csrc/kv_train.cu; KV_TRAIN_NATIVE=0 selects cuDNN): loss = cross-entropy(policy, move) + MSE(value, reward) - 0.01 * entropy (train.py:167-174), gradient
clipping at 1.0 and accumulation over 2 batches (train.py:183-190), bf16 autocast instead of fp16 + GradScaler.
With torch.distributed initialised (one process per GPU), gradients are averaged by DistributedDataParallel over
NCCL and every rank plays its own shard of the games.
"""
from __future__ import annotations

import logging
import os
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import train_ops as T
from .model import ChessNet
from .selfplay import SelfPlay, engine_for, filter_decisive_device

logger = logging.getLogger(__name__)
ENTROPY_COEF = 0.01


class TrainGraph(nn.Module):
    """Autograd graph of ai/model.py:51-77 over a ChessNet's parameters (train-mode BatchNorm).

    With `engine` given (the default on CUDA), the tower's 3x3 convolutions run forward, dgrad and wgrad on the
    tcgen05 kernels (train_ops.conv3x3_b200) and BatchNorm + ReLU (+ residual) on the fused NHWC kernels
    (train_ops.bn_relu_b200); otherwise everything is torch/cuDNN (the comparison arm)."""

    def __init__(self, net: ChessNet, engine=None):
        super().__init__()
        self.net = net
        self.engine = engine

    def _conv(self, m, h):
        if self.engine is not None and T.supported(m.in_channels, m.out_channels):
            return T.conv3x3_b200(h, m.weight, m.bias, self.engine)
        return m(h)

    def _bnrelu(self, bn, z, residual=None):
        if self.engine is not None and bn.training and T.bn_supported(bn.num_features):
            return T.bn_relu_b200(z, bn, self.engine, residual=residual)
        return F.relu(bn(z) if residual is None else bn(z) + residual)

    def forward(self, x):
        n = self.net
        h = n.conv1(x)
        if self.engine is not None:
            h = h.contiguous(memory_format=torch.channels_last)
        h = self._bnrelu(n.bn1, h)
        if n.arch[3]:
            h = self._bnrelu(n.bn2, self._conv(n.conv2, h))
        fused_block = (self.engine is not None and n.training and T.supported(n.arch[1], n.arch[1])
                       and T.bn_supported(n.arch[1]) and os.getenv("KV_TRAIN_BLOCK_FUSED", "1") != "0")
        for b in n.res_blocks:
            if fused_block:
                h = T.residual_block_b200(h, b, self.engine)       # one autograd node, skip gradient fused into dgrad
            else:
                t = self._bnrelu(b.bn1, self._conv(b.conv1, h))
                h = self._bnrelu(b.bn2, self._conv(b.conv2, t), residual=h)
        p = n.policy_fc(F.relu(n.policy_bn(n.policy_conv(h))).flatten(1))
        v = F.relu(n.value_bn(n.value_conv(h))).flatten(1)
        v = torch.tanh(n.value_fc2(F.relu(n.value_fc1(v))))
        return p, v


class ReplayData:
    """Packed device-resident training set with the reference's `extend` sink (scripts/train.py:560-561)."""

    def __init__(self, engine):
        self.eng = engine
        dev = engine.device
        self.lines = torch.zeros((0, 16), dtype=torch.int64, device=dev)
        self.move = torch.zeros(0, dtype=torch.int64, device=dev)
        self.reward = torch.zeros(0, dtype=torch.float32, device=dev)

    def __len__(self):
        return self.lines.shape[0]

    def extend_packed(self, lines, move, reward):
        self.lines = torch.cat([self.lines, lines])
        self.move = torch.cat([self.move, move.to(torch.int64)])
        self.reward = torch.cat([self.reward, reward])

    def extend(self, new_records):
        """Reference-format records [(np.float32 (12,8,8), move_index, reward)]: packed back to bitboards."""
        import numpy as np
        if not new_records:
            return
        planes = torch.from_numpy(np.stack([r[0] for r in new_records])).to(self.eng.device)
        w = (planes.reshape(len(new_records), 12, 64) != 0).to(torch.int64)
        bits = (w << torch.arange(64, device=w.device, dtype=torch.int64)).sum(-1)   # exact in two's complement
        lines = torch.zeros((len(new_records), 16), dtype=torch.int64, device=self.eng.device)
        lines[:, :12] = bits
        self.extend_packed(lines, torch.tensor([r[1] for r in new_records], device=self.eng.device),
                           torch.tensor([r[2] for r in new_records], dtype=torch.float32, device=self.eng.device))

    def batches(self, batch_size, generator=None):
        perm = torch.randperm(len(self), device=self.lines.device, generator=generator)
        for i in range(0, len(self), batch_size):
            idx = perm[i:i + batch_size]
            yield self.eng.encode(self.lines[idx].contiguous()), self.move[idx], self.reward[idx]


def _ddp_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def training_graph(net: ChessNet, engine):
    """The autograd graph of `net` (and, with torch.distributed initialised, its DistributedDataParallel wrapper), built
    once per network and reused by every train_epochs call: DDP broadcasts the parameters when it is constructed, so
    rebuilding it per call would re-send 100 MB per generation.
    Gradients travel as fp32 in 25 MB buckets.  Measured on 8 x B200 (profiles/r02_ddp_sweep_8gpu.txt): the NCCL
    all-reduce of the whole 95 MB gradient takes 0.38 ms (0.21 ms as bf16) against a 50 ms optimizer step, so the wire is
    not what the multi-GPU step pays for — the bf16 compression hook (KV_DDP_GRAD_BF16=1) made the step SLOWER
    (54.7 ms against 52.3 ms: its cast / copy kernels and hook calls cost more than the 0.17 ms they save), and one
    400 MB bucket (no overlap) slower still (57.2 ms)."""
    native = os.getenv("KV_TRAIN_NATIVE", "1") != "0"
    key = (id(engine) if native else None, _ddp_active())
    cached = getattr(net, "_kv_train_graph", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    graph = TrainGraph(net, engine=engine if native else None)
    if _ddp_active():
        on_gpu = next(net.parameters()).is_cuda
        graph = nn.parallel.DistributedDataParallel(graph, device_ids=[engine.index] if on_gpu else None,
                                                    gradient_as_bucket_view=True,
                                                    bucket_cap_mb=int(os.getenv("KV_DDP_BUCKET_MB", "25")))
        if os.getenv("KV_DDP_GRAD_BF16", "0") != "0":
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            graph.register_comm_hook(None, default_hooks.bf16_compress_hook)
    object.__setattr__(net, "_kv_train_graph", (key, graph))     # not a submodule: state_dict() keeps the reference's keys
    return graph


def train_epochs(net: ChessNet, optimizer, data: ReplayData, epochs: int, batch_size: int, accumulate_steps: int = 2):
    """scripts/train.py:126-196 without the logging side channels.  Returns the mean loss of the last epoch.
    Under DistributedDataParallel the gradient all-reduce runs once per optimizer step (micro-batches that do not step
    accumulate locally, `no_sync`), and a non-finite loss (train.py:176-178 skips the batch) is skipped by every rank
    together so that no rank waits for a gradient exchange the others never start."""
    import contextlib
    graph = training_graph(net, data.eng)
    net.train()
    last = float("nan")
    ddp = isinstance(graph, nn.parallel.DistributedDataParallel)
    for _ in range(epochs):
        tot, nb = 0.0, 0
        optimizer.zero_grad()
        batches = list(data.batches(batch_size))
        if ddp:
            # ranks hold different numbers of records (games end at different plies): every rank runs the same number
            # of steps, or the gradient all-reduce would wait for ever
            nmin = torch.tensor([len(batches)], device=data.eng.device)
            dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
            batches = batches[:int(nmin.item())]
        for i, (boards, moves, outcomes) in enumerate(batches):
            stepping = (i + 1) % accumulate_steps == 0 or i == len(batches) - 1
            with (graph.no_sync() if ddp and not stepping else contextlib.nullcontext()):
                with torch.autocast(boards.device.type, dtype=torch.bfloat16, enabled=boards.is_cuda):
                    pol, val = graph(boards)
                pol = pol.float()
                loss_policy = F.cross_entropy(pol, moves)
                loss_value = F.mse_loss(val.squeeze(1).float(), outcomes)
                logp = F.log_softmax(pol, dim=1)
                entropy = -(logp.exp() * logp).sum(dim=1).mean()
                loss = loss_policy + loss_value - ENTROPY_COEF * entropy
                ok = torch.isfinite(loss).to(torch.int32)
                if ddp:
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()):
                    (loss / accumulate_steps).backward()
                    tot += float(loss.item())
                    nb += 1
                elif ddp and stepping:
                    # the skipped batch was the stepping one: exchange what the earlier micro-batches accumulated
                    for p_ in net.parameters():
                        if p_.grad is not None:
                            dist.all_reduce(p_.grad, op=dist.ReduceOp.AVG)
            if stepping:
                torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
                optimizer.step()
                optimizer.zero_grad()
        last = tot / max(nb, 1)
    net.eval()                                             # train.py:444
    net.mark_weights_changed()                             # next forward / self-play re-uploads and re-folds the weights
    return last


def save_checkpoint(path: str, net: ChessNet, optimizer, epoch: int, loss=None):
    """The reference's checkpoint dictionary (scripts/train.py:207-212, :342-347): readable by
    ai/model_utils.py:10-29 and scripts/self_play.py:72-76 of an unmodified checkout."""
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    torch.save({"epoch": int(epoch), "model_state_dict": sd, "optimizer_state_dict": optimizer.state_dict(), "loss": loss},
               path)


def load_or_initialize_model(model_path, device, arch=None, lr: float = 1e-3):
    """ai/model_utils.py:10-29: (model, optimizer, start_epoch); a checkpoint missing the expected keys, or no file,
    gives a fresh model.  A bare state_dict (scripts/train.py:338-341) is accepted as well, DataParallel prefixes are
    stripped.  No DataParallel wrapper: multi-GPU runs are one process per GPU."""
    net = ChessNet(**(arch or {})).to(device)
    optimizer = torch.optim.Adam(net.parameters(), lr=lr)
    start_epoch = 0
    if model_path and os.path.exists(model_path):
        ck = torch.load(model_path, map_location="cpu")
        sd = ck.get("model_state_dict", ck) if isinstance(ck, dict) else ck
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        net.load_state_dict(sd)
        if isinstance(ck, dict) and "optimizer_state_dict" in ck:
            optimizer.load_state_dict(ck["optimizer_state_dict"])
            start_epoch = int(ck.get("epoch", 0))
    return net, optimizer, start_epoch


def build_cfg(**kw):
    """scripts/learn.py:99-149 without the Google-Drive / Stockfish parts; same environment variables."""
    cfg = SimpleNamespace(
        selfplay=SimpleNamespace(num_games=int(os.getenv("NUM_SELFPLAY_GAMES", "5")),
                                 max_moves=(int(os.environ["SELFPLAY_MAX_MOVES"]) if os.getenv("SELFPLAY_MAX_MOVES") else None),
                                 sims=int(os.getenv("KV_SIMS", "800"))),
        train=SimpleNamespace(epochs=int(os.getenv("TRAIN_EPOCHS", "2")), batch_size=int(os.getenv("BATCH_SIZE", "2048")),
                              lr=float(os.getenv("LR", "1e-3"))),
        num_iterations=int(os.getenv("NUM_ITERATIONS", "5")),
        device=torch.device("cuda"), model_path=None, arch={})
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def reinforcement_loop(cfg, net: ChessNet | None = None, data: ReplayData | None = None):
    """train -> self-play -> extend, cfg.num_iterations times.  Returns (net, data, history)."""
    eng = engine_for(cfg.device)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if net is None:
        net = ChessNet(**cfg.arch)
        if cfg.model_path and os.path.exists(cfg.model_path):
            ck = torch.load(cfg.model_path, map_location="cpu")
            net.load_state_dict(ck.get("model_state_dict", ck))
    net = net.to(cfg.device)
    optimizer = torch.optim.Adam(net.parameters(), lr=cfg.train.lr)
    data = data or ReplayData(eng)
    games = max(1, cfg.selfplay.num_games // world)
    history = []
    import time

    def now():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(1, cfg.num_iterations + 1):
        t0 = now()
        n_train = len(data)
        loss = train_epochs(net, optimizer, data, cfg.train.epochs, cfg.train.batch_size) if len(data) else float("nan")
        t1 = now()
        sp = SelfPlay(net.eval(), games, cfg.device, sims=cfg.selfplay.sims,
                      max_plies=cfg.selfplay.max_moves or 512, engine=eng,
                      inflight=getattr(cfg.selfplay, "inflight", None))
        if getattr(cfg.selfplay, "cache_log2", 0) and getattr(eng, "mcts_cache_log2", 0) != cfg.selfplay.cache_log2:
            eng.mcts_enable_cache(cfg.selfplay.cache_log2)       # (cleared by the weight commit above on later iterations)
        start = None
        if getattr(cfg.selfplay, "random_start_plies", 0):         # bench variant: diverse start positions
            start = eng.random_positions(games, cfg.selfplay.random_start_plies, 1234 + it * world + rank)
        t2 = now()
        st = sp.play(start, game_id_base=(it * world + rank) * games)
        t3 = now()
        lines, move, reward, _ = sp.records_device()
        if getattr(cfg.selfplay, "decisive_filter", True):
            # scripts/learn.py:186 calls generate_self_play_data, which keeps only the records of decisive games once
            # there are at least 10 of them (scripts/self_play.py:300-311); per rank, as the reference does per call
            lines, move, reward = filter_decisive_device(lines, move, reward)
        data.extend_packed(lines, move, reward)
        history.append(dict(iteration=it, loss=loss, records=int(lines.shape[0]), plies=st["plies"],
                            white=st["white_wins"], black=st["black_wins"], draws=st["draws"], evals=st["evals"],
                            train_positions=n_train * cfg.train.epochs, train_s=t1 - t0, weights_s=t2 - t1,
                            selfplay_s=t3 - t2, total_s=now() - t0))
        logger.info("iteration %d: loss %.4f, %d new records", it, loss, lines.shape[0])
    return net, data, history
