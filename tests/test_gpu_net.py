"""Policy/value network on the sm_100a kernels vs the fp32 reference graph (ai/model.py:51-77).
Tolerance (north_star): 2e-2 absolute on policy logits and value, bf16 kernels against fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers as H
from oracle import kv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def eng():
    from knightvision_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _net(seed=0, bnrand=False, **kw):
    from knightvision_b200.model import ChessNet
    torch.manual_seed(seed)
    net = ChessNet(**kw).eval()
    if bnrand:   # same recipe as oracle/gen_golden.py: non-trivial BN statistics so folding is exercised
        g = torch.Generator().manual_seed(1)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.05)
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
                m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.data.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    return net


def _lines(n, seed=3):
    lines = H.random_playout_positions(n_games=max(2, n // 40), max_plies=80, seed=seed)
    idx = np.random.default_rng(seed).permutation(len(lines))[:n]
    return lines[idx]


def test_layers_against_torch(eng):
    """stem, conv2 and the first residual block, layer by layer (localises a GEMM/TMA/descriptor bug)."""
    from knightvision_b200.engine import lines_to_device
    from knightvision_b200.model import fp32_reference_forward  # noqa: F401
    net = _net(bnrand=True).attach(eng, max_batch=64)
    lines = _lines(37)
    d = lines_to_device(lines, eng.device)
    x = torch.from_numpy(O.encode(lines)).cuda()
    netc = net.cuda()

    def bn(m, t):
        return F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)
    with torch.no_grad():
        a0 = F.relu(bn(netc.bn1, netc.conv1(x)))
        g0 = eng.net_forward_partial(d, 0).float().permute(0, 3, 1, 2)
        assert (g0 - a0).abs().max().item() < 1e-2
        a1 = F.relu(bn(netc.bn2, netc.conv2(a0)))
        g1 = eng.net_forward_partial(d, 1).float().permute(0, 3, 1, 2)
        assert (g1 - a1).abs().max().item() < 2e-2, (g1 - a1).abs().max().item()
        b = netc.res_blocks[0]
        t = F.relu(bn(b.bn1, b.conv1(a1)))
        g2 = eng.net_forward_partial(d, 2).float().permute(0, 3, 1, 2)
        assert (g2 - t).abs().max().item() < 3e-2, (g2 - t).abs().max().item()
        a2 = F.relu(bn(b.bn2, b.conv2(t)) + a1)
        g3 = eng.net_forward_partial(d, 3).float().permute(0, 3, 1, 2)
        assert (g3 - a2).abs().max().item() < 4e-2, (g3 - a2).abs().max().item()


@pytest.mark.parametrize("variant", ["init", "bnrand"])
def test_forward_matches_reference_golden(eng, variant):
    """Golden outputs of the UNMODIFIED reference ChessNet (fp32, CPU) on 8 playout positions."""
    from knightvision_b200.engine import lines_to_device
    g = np.load(H.GOLDEN + "/net.npz")
    net = _net(bnrand=(variant == "bnrand")).attach(eng, max_batch=64)
    assert list(g["keys"]) == list(net.state_dict().keys())
    pol, val = net.forward_lines(lines_to_device(g["lines"], eng.device))
    torch.cuda.synchronize()
    dp = np.abs(pol.cpu().numpy() - g["policy_" + variant]).max()
    dv = np.abs(val.cpu().numpy() - g["value_" + variant]).max()
    assert dp < TOL and dv < TOL, (dp, dv)
    # the planes entry point (ChessNet.forward drop-in) gives the same numbers
    x = torch.from_numpy(O.encode(g["lines"])).cuda()
    pol2, val2 = net(x)
    assert torch.equal(pol2, pol) and torch.equal(val2, val)


def _wide_errors(g, prefix, pol, val):
    pol = pol.cpu().numpy()
    n = pol.shape[0]
    cols = g["cols"][:n].astype(np.int64)
    d_cols = np.abs(np.take_along_axis(pol, cols, axis=1) - g[prefix + "_cols"]).max()
    d_top = np.abs(np.take_along_axis(pol, g[prefix + "_top_idx"].astype(np.int64), axis=1) - g[prefix + "_top_val"]).max()
    d_val = np.abs(val.cpu().numpy() - g[prefix + "_value"]).max()
    return float(d_cols), float(d_top), float(d_val)


@pytest.mark.parametrize("variant", ["init", "bnrand"])
def test_forward_matches_reference_golden_256_positions(eng, variant, capsys):
    """tests/golden/net_wide.npz: the UNMODIFIED reference ChessNet in true fp32 on the CPU, 256 positions (208 playout
    + 48 synthetic), 256 random policy columns + the 8 largest logits + the value per position.  Prints the observed
    maximum absolute error next to the 2e-2 bound."""
    from knightvision_b200.engine import lines_to_device
    g = np.load(H.GOLDEN + "/net_wide.npz")
    net = _net(bnrand=(variant == "bnrand")).attach(eng, max_batch=256)
    pol, val = net.forward_lines(lines_to_device(g["lines"], eng.device))
    d_cols, d_top, d_val = _wide_errors(g, "ref_" + variant, pol, val)
    with capsys.disabled():
        print(f"\n[net parity, reference net, {variant}] max |bf16 kernel - fp32 CPU reference| over 256 positions: "
              f"policy {max(d_cols, d_top):.3e}, value {d_val:.3e} (bound {TOL:.0e})")
    assert max(d_cols, d_top) < TOL and d_val < TOL, (d_cols, d_top, d_val)


def test_tower_20x256_matches_cpu_fp32_golden(eng, capsys):
    """BASELINE config 5's 20 x 256 tower against a plain torch CPU fp32 run of the same graph (128 positions)."""
    from knightvision_b200.engine import lines_to_device
    g = np.load(H.GOLDEN + "/net_wide.npz")
    n = g["t20_value"].shape[0]
    net = _net(seed=2, bnrand=True, stem=256, tower=256, blocks=20, conv2=False).attach(eng, max_batch=n)
    pol, val = net.forward_lines(lines_to_device(g["lines"][:n], eng.device))
    d_cols, d_top, d_val = _wide_errors(g, "t20", pol, val)
    with capsys.disabled():
        print(f"\n[net parity, 20x256 tower] max |bf16 kernel - fp32 CPU| over {n} positions: "
              f"policy {max(d_cols, d_top):.3e}, value {d_val:.3e} (bound {TOL:.0e})")
    assert max(d_cols, d_top) < TOL and d_val < TOL, (d_cols, d_top, d_val)


def test_torch_references_are_true_fp32(eng):
    """The torch graphs these tests compare against must not run TF32 (conftest switches it off globally and
    fp32_reference_forward does so locally): a GPU fp32 run agrees with the CPU fp32 golden to fp32 rounding."""
    from knightvision_b200.model import fp32_reference_forward
    assert not torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32
    g = np.load(H.GOLDEN + "/net_wide.npz")
    net = _net(bnrand=True).cuda()
    x = torch.from_numpy(O.encode(g["lines"][:32])).cuda()
    torch.backends.cudnn.allow_tf32 = True          # the call must be immune to the global switch
    try:
        with torch.no_grad():
            rp, rv = fp32_reference_forward(net, x)
    finally:
        torch.backends.cudnn.allow_tf32 = False
    sub = {k: g[k][:32] for k in ("cols", "ref_bnrand_cols", "ref_bnrand_top_idx", "ref_bnrand_top_val", "ref_bnrand_value")}
    d_cols, d_top, d_val = _wide_errors(sub, "ref_bnrand", rp, rv)
    assert max(d_cols, d_top, d_val) < 2e-4, (d_cols, d_top, d_val)


def test_forward_batch_odd_sizes_and_determinism(eng):
    from knightvision_b200.engine import lines_to_device
    from knightvision_b200.model import fp32_reference_forward
    net = _net(seed=5, bnrand=True).attach(eng, max_batch=700)
    lines = _lines(611, seed=9)
    d = lines_to_device(lines, eng.device)
    pol, val = net.forward_lines(d)
    pol_b, val_b = net.forward_lines(d)
    assert torch.equal(pol, pol_b) and torch.equal(val, val_b)          # run-to-run deterministic
    p1, v1 = net.forward_lines(d[:1])
    assert torch.equal(p1[0], pol[0]) and torch.equal(v1[0], val[0])     # independent of batch composition
    with torch.no_grad():
        rp, rv = fp32_reference_forward(net.cuda(), torch.from_numpy(O.encode(lines)).cuda())
    assert (pol - rp).abs().max().item() < TOL and (val - rv).abs().max().item() < TOL


def test_rejects_training_mode_and_non_onehot(eng):
    from knightvision_b200 import _native as N
    net = _net().attach(eng, max_batch=8)
    with pytest.raises(N.KVError):
        net.train()(torch.zeros(1, 12, 8, 8).cuda())
    net.eval()
    with pytest.raises(N.KVError):
        net(torch.full((1, 12, 8, 8), 0.5).cuda())


def test_tower_20x256_variant(eng):
    """BASELINE config 5's 20-block x 256-channel tower through the same kernels."""
    from knightvision_b200.engine import lines_to_device
    from knightvision_b200.model import fp32_reference_forward
    net = _net(seed=2, bnrand=True, stem=256, tower=256, blocks=20, conv2=False).attach(eng, max_batch=64)
    lines = _lines(33, seed=4)
    pol, val = net.forward_lines(lines_to_device(lines, eng.device))
    with torch.no_grad():
        rp, rv = fp32_reference_forward(net.cuda(), torch.from_numpy(O.encode(lines)).cuda())
    assert (pol - rp).abs().max().item() < TOL and (val - rv).abs().max().item() < TOL


def test_cta_pair_kernel_matches_single_cta_kernel(eng):
    """cta_group::2 (CTA pairs, M = 256) and cta_group::1 tower kernels: bit-identical activations and outputs."""
    from knightvision_b200.engine import lines_to_device
    net = _net(seed=7, bnrand=True).attach(eng, max_batch=300)
    lines = _lines(261, seed=12)
    d = lines_to_device(lines, eng.device)
    outs = {}
    eng.net_set_tower_fused(1)      # the default halo-operand kernel accumulates the taps in another order
    for mode in (1, 2):
        eng.net_set_conv_mode(mode)
        outs[mode] = (eng.net_forward_partial(d, 1).clone(), eng.net_forward_partial(d, 3).clone(), net.forward_lines(d))
    eng.net_set_conv_mode(2)
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])
    assert torch.equal(outs[1][2][0], outs[2][2][0]) and torch.equal(outs[1][2][1], outs[2][2][1])


@pytest.mark.parametrize("arch,n", [("ref", 1237), ("ref", 9), ("20x256", 333)])
def test_whole_tower_launch_matches_per_layer_launches(eng, arch, n):
    """tower_umma2_kernel (all layers in one launch, tiles scheduled by their dependencies) against one launch per
    layer: same tiles, same arithmetic -> bit-identical activations and outputs; repeated runs agree (a missed
    dependency would show up as run-to-run differences)."""
    from knightvision_b200.engine import lines_to_device
    kw = {} if arch == "ref" else dict(stem=256, tower=256, blocks=20, conv2=False)
    net = _net(seed=11, bnrand=True, **kw).attach(eng, max_batch=n + 3)
    lines = _lines(n, seed=21)
    d = lines_to_device(lines, eng.device)
    eng.net_set_tower_fused(False)
    ref_mid = eng.net_forward_partial(d, 4).clone()
    ref_pol, ref_val = (t.clone() for t in net.forward_lines(d))
    eng.net_set_tower_fused(1)
    for _ in range(4):
        assert torch.equal(eng.net_forward_partial(d, 4), ref_mid)
        pol, val = net.forward_lines(d)
        assert torch.equal(pol, ref_pol) and torch.equal(val, ref_val)


@pytest.mark.parametrize("arch,n", [("ref", 1237), ("ref", 6), ("20x256", 333)])
def test_halo_operand_tower(eng, arch, n):
    """Whole-tower launch with the halo activation operand (3 TMA fetches per tile and channel block instead of 9, row
    shifts as descriptor offsets): same network within the north-star tolerance, equal to the 9-fetch kernel up to fp32
    accumulation order, deterministic."""
    from knightvision_b200.engine import lines_to_device
    from knightvision_b200.model import fp32_reference_forward
    kw = {} if arch == "ref" else dict(stem=256, tower=256, blocks=20, conv2=False)
    net = _net(seed=13, bnrand=True, **kw).attach(eng, max_batch=n + 2)
    lines = _lines(n, seed=23)
    d = lines_to_device(lines, eng.device)
    eng.net_set_tower_fused(1)
    mid1 = eng.net_forward_partial(d, 3).float()
    pol1, val1 = (t.clone() for t in net.forward_lines(d))
    eng.net_set_tower_fused(2)
    mid2 = eng.net_forward_partial(d, 3).float()
    pol2, val2 = (t.clone() for t in net.forward_lines(d))
    pol3, val3 = net.forward_lines(d)
    assert torch.equal(pol2, pol3) and torch.equal(val2, val3)
    # bf16 activations: a different fp32 accumulation order flips a rounding here and there (1 bf16 ulp = 2^-8 relative)
    assert ((mid1 - mid2).abs() <= 0.02 * mid1.abs().clamp(min=1.0)).all()
    assert (pol1 - pol2).abs().max().item() < 5e-3 and (val1 - val2).abs().max().item() < 5e-3
    with torch.no_grad():
        rp, rv = fp32_reference_forward(net.cuda(), torch.from_numpy(O.encode(lines)).cuda())
    assert (pol2 - rp).abs().max().item() < TOL and (val2 - rv).abs().max().item() < TOL


@pytest.mark.parametrize("arch,n", [("ref", 1237), ("ref", 5), ("ref", 24), ("20x256", 333)])
def test_weight_multicast_clusters_match_pairs(eng, arch, n):
    """Mode 3 (4-CTA clusters, weight tiles multicast to two CTA pairs) computes every tile exactly like mode 2:
    bit-identical activations and outputs, odd numbers of board tiles included; repeated runs agree."""
    from knightvision_b200.engine import lines_to_device
    kw = {} if arch == "ref" else dict(stem=256, tower=256, blocks=20, conv2=False)
    net = _net(seed=17, bnrand=True, **kw).attach(eng, max_batch=n + 2)
    if eng.net_tower_clusters4() < 8:
        pytest.skip("4-CTA clusters of the tower kernel are not co-resident on this GPU")
    lines = _lines(n, seed=29)
    d = lines_to_device(lines, eng.device)
    eng.net_set_tower_fused(2)
    mid2 = eng.net_forward_partial(d, 3).clone()
    pol2, val2 = (t.clone() for t in net.forward_lines(d))
    for mode in (3, 4):      # 4 = hybrid: the clusters on the SMs they can cover, CTA pairs on the rest, concurrently
        eng.net_set_tower_fused(mode)
        for _ in range(3):
            assert torch.equal(eng.net_forward_partial(d, 3), mid2)
            pol3, val3 = net.forward_lines(d)
            assert torch.equal(pol3, pol2) and torch.equal(val3, val2)
    eng.net_set_tower_fused(2)
