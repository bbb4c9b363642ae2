"""Run one of the reference's scripts UNCHANGED with the B200 modules substituted for `utils.networks` and
`utils.loss_functions` (and an fvcore stand-in if fvcore/yacs are not installed).

    python -m multimodal_siamese_cd_b200.launcher REF_ROOT SCRIPT [script args...]
    e.g.  ... launcher /path/to/multimodal_siamese_cd train_supervised.py -c baseline_siamese -o OUT -d DATA

The reference tree is only read. Under torchrun (WORLD_SIZE > 1) the process group is initialised and the reference's
nn.DataParallel semantics are enabled in "gather" mode (parallel.enable_data_parallel(mode="gather")): every rank's
unchanged DataLoader yields the same full batch (same seed), the network runs this rank's DataParallel chunk and
returns the gathered full-batch logits, the script's loss code runs on the global batch as written, and gradients are
SUM-reduced in buckets overlapped with backward. Side effects of the scripts happen on rank 0 only: checkpoints
(networks.save_checkpoint) are written by rank 0, wandb is disabled on the other ranks.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys


def substitute_modules(ref_root: str) -> None:
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from .config import install_fvcore_stub
    install_fvcore_stub()
    from . import loss_functions, networks
    utils_pkg = importlib.import_module("utils")          # the reference's package (namespace or regular)
    sys.modules["utils.networks"] = networks
    sys.modules["utils.loss_functions"] = loss_functions
    utils_pkg.networks = networks
    utils_pkg.loss_functions = loss_functions
    # the evaluation loop's metric (utils/evaluation.py:12,23): same class name and interface, one kernel per sample.
    # Only the class is swapped; the module's numpy helpers stay the reference's.
    try:
        ref_metrics = importlib.import_module("utils.metrics")
        from . import metrics as b200_metrics
        ref_metrics.MultiThresholdMetric = b200_metrics.MultiThresholdMetric
    except Exception:  # noqa: BLE001  (a reference tree without utils/metrics.py still trains)
        pass


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(__doc__)
    ref_root, script, rest = os.path.abspath(argv[0]), argv[1], argv[2:]
    substitute_modules(ref_root)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist

        from . import parallel
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        parallel.enable_data_parallel(mode="gather")
        if dist.get_rank() != 0:
            os.environ["WANDB_MODE"] = "disabled"
        from . import networks
        _save = networks.save_checkpoint

        def save_checkpoint_rank0(*args, **kwargs):
            if dist.get_rank() == 0:
                _save(*args, **kwargs)
            dist.barrier()                                 # nobody reads a checkpoint that is still being written

        networks.save_checkpoint = save_checkpoint_rank0
    os.chdir(ref_root)                                     # the scripts read configs/<name>.yaml relative to cwd
    sys.argv = [script, *rest]
    runpy.run_path(os.path.join(ref_root, script), run_name="__main__")


if __name__ == "__main__":
    main()
