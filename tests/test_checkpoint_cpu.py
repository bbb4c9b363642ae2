"""N4: checkpoint files round-trip between the reference and the drop-in modules.
tests/golden/ckpt/ref_checkpoint_siamese_tiny.pt was written by the UNMODIFIED reference's save_checkpoint
(utils/networks.py:30-38) after one AdamW step (oracle/make_checkpoint_golden.py). Pure CPU: the drop-in modules are
ordinary nn.Modules until forward() is called."""
from pathlib import Path

import torch

from multimodal_siamese_cd_b200 import networks
from multimodal_siamese_cd_b200.config import synthetic_cfg

FIX = Path(__file__).parent / "golden" / "ckpt" / "ref_checkpoint_siamese_tiny.pt"


def _cfg(tmp_path):
    cfg = synthetic_cfg("siameseunet", in_channels=4, topology=(8, 16))
    cfg.PATHS.OUTPUT = str(tmp_path)
    cfg.NAME = "tiny"
    return cfg


def test_reference_checkpoint_loads_as_is(tmp_path):
    ref = torch.load(FIX, map_location="cpu")
    assert set(ref) == {"step", "network", "optimizer"} and ref["step"] == 17
    cfg = _cfg(tmp_path)
    net, opt, step = networks.load_checkpoint(3, cfg, torch.device("cpu"), net_file=FIX)
    assert step == 17
    sd = net.state_dict()
    assert list(sd) == list(ref["network"]), "state_dict keys (and their order) must equal the reference's"
    for k, v in ref["network"].items():
        assert torch.equal(sd[k], v), k
    osd = opt.state_dict()
    assert osd["param_groups"][0]["lr"] == cfg.TRAINER.LR and osd["param_groups"][0]["weight_decay"] == 0.01
    assert len(osd["state"]) == len(ref["optimizer"]["state"])
    for i, st in ref["optimizer"]["state"].items():
        for name in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(osd["state"][i][name], st[name]), (i, name)
        assert float(osd["state"][i]["step"]) == float(st["step"]) == 1.0
    # outc_sem_change does not exist in a siamese net; every parameter took part in the step
    assert len(osd["state"]) == sum(1 for _ in net.parameters())


def test_dropin_checkpoint_has_the_reference_layout(tmp_path):
    ref = torch.load(FIX, map_location="cpu")
    cfg = _cfg(tmp_path)
    net, opt, step = networks.load_checkpoint(3, cfg, torch.device("cpu"), net_file=FIX)
    networks.save_checkpoint(net, opt, 4, step + 5, cfg)                    # utils/networks.py:30-38 naming
    written = Path(cfg.PATHS.OUTPUT) / "networks" / "tiny_checkpoint4.pt"
    assert written.exists()
    ours = torch.load(written, map_location="cpu")
    assert ours["step"] == 22 and set(ours) == set(ref)
    assert list(ours["network"]) == list(ref["network"])
    assert all(torch.equal(ours["network"][k], ref["network"][k]) for k in ref["network"])
    assert ours["optimizer"]["param_groups"][0]["params"] == ref["optimizer"]["param_groups"][0]["params"]
    # and the file written by the drop-in loads back through the default path (epoch -> file name)
    net2, opt2, step2 = networks.load_checkpoint(4, cfg, torch.device("cpu"))
    assert step2 == 22 and all(torch.equal(a, b) for a, b in zip(net2.state_dict().values(), net.state_dict().values()))
