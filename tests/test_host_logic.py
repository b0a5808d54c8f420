"""Host-side logic that needs no GPU: checkpoint format bridge."""
def test_checkpoint_format_round_trip(tmp_path):
    """scripts/train.py:207-212 dictionary: written by learn.save_checkpoint, read by learn.load_or_initialize_model, and
    its model_state_dict has the reference's 104 tensors — CPU only, no kernels involved."""
    import torch
    from knightvision_b200 import learn as LR
    from knightvision_b200.model import ChessNet
    torch.manual_seed(0)
    net = ChessNet()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    p = str(tmp_path / "best_model.pth")
    LR.save_checkpoint(p, net, opt, epoch=3, loss=1.25)
    ck = torch.load(p, map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"} and ck["epoch"] == 3
    keys = list(ck["model_state_dict"])
    assert len(keys) == 104 and keys[0] == "conv1.weight" and "res_blocks.4.bn2.running_var" in keys
    assert sum(v.numel() for k, v in ck["model_state_dict"].items() if not k.endswith("num_batches_tracked")
               and "running" not in k) == 25381642                          # the reference's parameter count
    net2, opt2, ep = LR.load_or_initialize_model(p, "cpu")
    assert ep == 3 and all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    torch.save({"module." + k: v for k, v in net.state_dict().items()}, str(tmp_path / "bare.pth"))   # DataParallel-style
    net3, _, ep3 = LR.load_or_initialize_model(str(tmp_path / "bare.pth"), "cpu")
    assert ep3 == 0 and torch.equal(net3.conv2.weight, net.conv2.weight)
    _, _, ep4 = LR.load_or_initialize_model(str(tmp_path / "missing.pth"), "cpu")
    assert ep4 == 0

