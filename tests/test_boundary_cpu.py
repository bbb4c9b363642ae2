"""CPU: the drop-in boundary — C-ABI symbols, loud failure without CUDA, config shim, planner dry-run, gradient
arena layout and the bucket plan of the data-parallel backward. No kernel is launched here."""
from __future__ import annotations

import re
from pathlib import Path

import pytest
import torch

from multimodal_siamese_cd_b200 import _lib, loss_functions, networks, parallel
from multimodal_siamese_cd_b200.config import CfgNode, load_cfg, synthetic_cfg
from multimodal_siamese_cd_b200.engine import StepEngine

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "b200cd.h").read_text()
    declared = set(re.findall(r"\b(b200cd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200cd.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.b200cd_abi_version() == 2
    # pure host helpers may be called without a GPU
    assert lib.b200cd_conv_gemm_tiles(256, 256) == 512
    assert lib.b200cd_wgrad_tiles(2, 32, 32) == 32


def test_no_cpu_fallback():
    net = networks.create_network(synthetic_cfg("siameseunet", in_channels=4, topology=(64, 128)))
    x = torch.rand(1, 4, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss_functions.get_criterion("PowerJaccardLoss")(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))
    assert issubclass(_lib.B200CDError, RuntimeError)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.B200CDError):
            _lib.init(0)  # fails loudly (no driver / wrong architecture), never silently


def test_get_criterion_surface():
    for name in ("PowerJaccardLoss", "BCEWithLogitsLoss", "SoftDiceLoss", "SoftDiceSquaredSumLoss", "SoftDiceBalancedLoss",
                 "MeanSquareErrorLoss", "IoULoss", "DiceLikeLoss", "L2"):
        assert callable(loss_functions.get_criterion(name))
    with pytest.raises(Exception, match="unknown loss"):
        loss_functions.get_criterion("nope")
    with pytest.raises(Exception, match="Unknown network"):
        networks.create_network(synthetic_cfg("resnet"))


def test_cfg_shim_yaml_inheritance(tmp_path):
    (tmp_path / "base.yaml").write_text("SEED: 7\nTRAINER:\n  LR: 1e-4\n  BATCH_SIZE: 8\nMODEL:\n  TYPE: 'unet'\n  IN_CHANNELS: 3\n"
                                        "  TOPOLOGY: [64, 128, 256, 512, ]\nDATALOADER:\n  S1_BANDS: [0, 1]\n  S2_BANDS: [2, 1, 0, 3]\n")
    (tmp_path / "child.yaml").write_text('_BASE_: "base.yaml"\nDEBUG: True\nMODEL:\n  TYPE: \'siameseunet\'\n  IN_CHANNELS: 4\n')
    cfg = load_cfg(tmp_path / "child.yaml", ["TRAINER.BATCH_SIZE", "16", "MODEL.NEW_KEY", "abc"])
    assert cfg.MODEL.TYPE == "siameseunet" and cfg.MODEL.IN_CHANNELS == 4 and cfg.MODEL.TOPOLOGY == [64, 128, 256, 512]
    assert type(cfg.TRAINER.LR) is float and cfg.TRAINER.LR == 1e-4          # PyYAML reads '1e-4' as a string
    assert cfg.TRAINER.BATCH_SIZE == 16 and cfg.MODEL.NEW_KEY == "abc" and cfg.DEBUG is True and cfg.NAME == "child"
    assert isinstance(cfg.MODEL, CfgNode) and cfg.clone().MODEL.TYPE == "siameseunet"


@pytest.mark.parametrize("mtype,cin,n_params", [("unet", 6, 14794113), ("siameseunet", 4, 14789505),
                                                ("dualstreamunet", 6, 29581313), ("dtsiameseunet", 6, 20168709),
                                                ("whatevernet", 6, 29577987), ("whatevernet2", 6, 29581443)])
def test_parameter_counts_and_plan(mtype, cin, n_params):
    """Parameter counts of SURVEY App. A.2 and a dry run of the planner (meta device: no kernels, no memory)."""
    net = networks.create_network(synthetic_cfg(mtype, in_channels=cin))
    assert sum(p.numel() for p in net.parameters()) == n_params
    assert all(k.startswith("module.") for k in net.state_dict())
    eng = StepEngine(net.module, 2, 64, 64, True, torch.device("meta"))
    g = eng.grads
    # every parameter except outc_sem_change has exactly one slot; slots tile the flat buffer in write order
    slots = sorted((g.offsets[n], g.offsets[n] + p.numel()) for n, p in g.params if n not in g.skip)
    assert all(a1 >= b0 for (_, b0), (a1, _) in zip(slots, slots[1:]))
    assert {n for n, _ in g.params if n in g.skip} == {n for n, _ in g.params if n.startswith("outc_sem_change")}
    assert len(eng.bwd_marks) == len(eng.bwd_ops) and eng.bwd_marks == sorted(eng.bwd_marks)
    assert eng.bwd_marks[-1] == g.flat.numel()
    for st in eng.stages:
        assert 1 <= len(st.srcs) <= 3, st.name


@pytest.mark.parametrize("mtype,two", [("unet", False), ("siameseunet", False), ("dualstreamunet", True),
                                       ("dtsiameseunet", True), ("whatevernet", True), ("whatevernet2", True)])
def test_branch_plan_respects_dependencies(mtype, two):
    """Two-stream plans (engine._run_ops): one branch code per op; a branch-1 op may only consume main-stream results
    that were queued before its branch was forked, and main-stream ops that consume branch-1 results sit behind a join."""
    net = networks.create_network(synthetic_cfg(mtype, in_channels=6 if mtype != "siameseunet" else 4))
    eng = StepEngine(net.module, 2, 64, 64, True, torch.device("meta"))
    assert len(eng.fwd_branch) == len(eng.fwd_ops) and len(eng.bwd_branch) == len(eng.bwd_ops)
    assert set(eng.fwd_branch) <= {0, 1, 10, -1} and set(eng.bwd_branch) <= {0, 1, -10, -1}
    assert any(b in (1, 10) for b in eng.fwd_branch) == two
    assert eng.fwd_branch[-1] == -1 and eng.bwd_branch[0] == -1            # heads: last forward, first backward
    if not two:
        return
    if mtype == "dtsiameseunet":
        # forward: encoder (0) ... fork (10) ... decoder_sem (1) ... decoder_change (0) ... heads (-1)
        i10 = eng.fwd_branch.index(10)
        assert all(b == 0 for b in eng.fwd_branch[:i10]), "everything before the fork is the shared encoder"
        assert eng.fwd_branch.count(10) == 1
        after = [b for b in eng.fwd_branch[i10 + 1:] if b != -1]
        first0 = after.index(0)
        assert all(b == 1 for b in after[:first0]) and all(b == 0 for b in after[first0:]), \
            "decoder_sem's ops directly follow the fork, decoder_change's come after them"
        # backward: heads, the two decoders, then ONE join in front of the first encoder op, encoder on the main stream
        j = eng.bwd_branch.index(-10)
        assert eng.bwd_branch.count(-10) == 1 and all(b == 0 for b in eng.bwd_branch[j + 1:])
        n_enc = sum(1 for st in eng.stages if st.name.startswith("s.inc") or st.name.startswith("s.down"))
        assert len(eng.bwd_branch) - j == n_enc
    else:
        # two trunks: no cross-trunk dependency before the heads, so no fork/join codes at all
        assert 10 not in eng.fwd_branch and -10 not in eng.bwd_branch
        assert sum(b == 1 for b in eng.fwd_branch) == sum(b == 0 for b in eng.fwd_branch)


def test_bucket_plan_covers_gradient_buffer_once():
    net = networks.create_network(synthetic_cfg("dtsiameseunet", in_channels=6)).module
    eng = StepEngine(net, 2, 64, 64, True, torch.device("meta"))
    plan = eng.plan_buckets(4)
    assert plan[0][0] == 0 and plan[-1][1] == len(eng.bwd_ops)
    assert plan[0][2] == 0 and plan[-1][3] == eng.grads.flat.numel()
    for (o0, o1, g0, g1), (p0, p1, h0, h1) in zip(plan, plan[1:]):
        assert o1 == p0 and g1 == h0 and o1 > o0 and g1 > g0


def test_shard_rows_matches_dataparallel_scatter():
    assert [parallel.shard_rows(8, r, 2) for r in range(2)] == [slice(0, 4), slice(4, 8)]
    assert [parallel.shard_rows(10, r, 4) for r in range(4)] == [slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 10)]


def test_launcher_substitutes_reference_modules(tmp_path, monkeypatch):
    """A script written like the reference's drivers (`from utils import networks, loss_functions`) gets the B200
    modules; the (fake) reference tree is not modified."""
    import sys

    from multimodal_siamese_cd_b200 import launcher
    ref = tmp_path / "ref"
    (ref / "utils").mkdir(parents=True)
    (ref / "utils" / "__init__.py").write_text("")
    (ref / "utils" / "networks.py").write_text("raise ImportError('the reference module must not be imported')\n")
    (ref / "utils" / "loss_functions.py").write_text("raise ImportError('the reference module must not be imported')\n")
    (ref / "utils" / "experiment_manager.py").write_text("from fvcore.common.config import CfgNode as _CfgNode\n")
    (ref / "driver.py").write_text(
        "import sys\nfrom utils import networks, loss_functions, experiment_manager\n"
        "open(sys.argv[1], 'w').write(networks.__name__ + ' ' + loss_functions.__name__ + ' ' + "
        "experiment_manager._CfgNode.__module__)\n")
    out = tmp_path / "out.txt"
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.networks", "utils.loss_functions", "utils.experiment_manager")}
    monkeypatch.chdir(tmp_path)
    try:
        launcher.main([str(ref), "driver.py", str(out)])
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        if str(ref) in sys.path:
            sys.path.remove(str(ref))
    got = out.read_text().split()
    assert got[0] == "multimodal_siamese_cd_b200.networks" and got[1] == "multimodal_siamese_cd_b200.loss_functions"
    assert got[2] in ("multimodal_siamese_cd_b200.config", "fvcore.common.config")


def test_documented_knobs_exist_in_code():
    """Every B200CD_* environment variable named in INTEGRATION.md's knob table is read somewhere in the product (and
    every one the product reads is documented)."""
    import re
    root = Path(__file__).resolve().parent.parent
    doc = (root / "INTEGRATION.md").read_text()
    table = doc[doc.index("## Tuning knobs"):]
    documented = set(re.findall(r"`(B200CD_[A-Z0-9_]+)\*?`", table))
    code = ""
    for f in list((root / "multimodal_siamese_cd_b200").rglob("*.py")) + list((root / "multimodal_siamese_cd_b200" / "csrc").glob("*.cu")) \
            + list((root / "multimodal_siamese_cd_b200" / "csrc").glob("*.h")) + [root / "bench.py"]:
        code += f.read_text()
    read = set(re.findall(r"(?:getenv|environ\.get|environ\[)\(?\s*\"(B200CD_[A-Z0-9_]+)\"", code))
    debug = {k for k in read if k.startswith("B200CD_DEBUG_")}
    assert documented - {"B200CD_DEBUG_"} <= read | {"B200CD_NVCC_EXTRA"}, documented - read
    assert read - debug <= documented, read - debug - documented



def test_tile_table_is_well_formed_and_reaches_the_flags(tmp_path):
    """tuning.TABLE: every entry names a known variant, a tile that divides N and differs from nothing the kernels cannot
    run; ops._conv_flags encodes it in bits 5..6 (include/b200cd.h); B200CD_TILE_TABLE-style JSON round-trips; the
    committed table equals profiles/r02_tile_table_v3.json (the measured one)."""
    import json
    from multimodal_siamese_cd_b200 import ops, tuning
    assert tuning.TABLE, "the measured table is committed"
    for key, bn in tuning.TABLE.items():
        variant, mode, out_mode, n, H, W, ka, N, prec = key
        assert variant in ("stats", "plain", "bnbwd", "affine") and mode in (0, 1, 2) and out_mode in (0, 1)
        assert bn in (64, 128, 256) and N % bn == 0 and ka % 64 == 0 and isinstance(prec, bool)
        assert not (variant == "bnbwd" and prec), "the fused BatchNorm-backward dgrad is a bf16-storage kernel"
        assert tuning.tile(*key) == bn
        flags = ops._conv_flags(mode, out_mode, N, ka, None, None, None, 1 if variant in ("stats", "bnbwd") else 0, bn)
        assert flags & 4 and ((flags >> 5) & 3) == {64: 1, 128: 2, 256: 3}[bn]
    assert ops._conv_flags(0, 0, 256, 64, None, None, None, 0, None) == 4          # no entry: the library's own rule
    root = Path(__file__).resolve().parent.parent
    measured = json.loads((root / "profiles" / "r02_tile_table_v3.json").read_text())
    assert {tuple(e[:8]) + (bool(e[8]),): e[9] for e in measured} == tuning.TABLE
    saved = dict(tuning.TABLE)
    try:
        f = tmp_path / "t.json"
        f.write_text(json.dumps([["plain", 0, 0, 2, 32, 32, 64, 256, 0, 128]]))
        tuning.load_table(str(f))
        assert tuning.TABLE == {("plain", 0, 0, 2, 32, 32, 64, 256, False): 128}
        old, tuning.ENABLED = tuning.ENABLED, False
        assert tuning.tile("plain", 0, 0, 2, 32, 32, 64, 256, False) is None         # B200CD_TUNED_TILES=0
        tuning.ENABLED = old
    finally:
        tuning.TABLE.clear()
        tuning.TABLE.update(saved)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): ONE JSON line on stdout with the base contract's keys,
    `impl: reference`, a cpu_baseline describing the run and an e2e object equal to the line's own value — measured with
    the reference's own modules when the staged tree exists (kind "reference"), else the oracle port."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1",
                        "--config", "siamese"], capture_output=True, text=True, timeout=900, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "patch-pairs/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
