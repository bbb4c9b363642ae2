"""TEST-ONLY stand-in for `wandb`: records what the reference's scripts log (train_supervised.py:93-99,
utils/evaluation.py:35-40) as JSON lines in $B200CD_STUB_WANDB_LOG instead of talking to a server. Injected through
PYTHONPATH by tests/test_reference_scripts.py only."""
import json
import os

config = None


def init(*args, **kwargs):
    return None


def log(data, *args, **kwargs):
    path = os.environ.get("B200CD_STUB_WANDB_LOG")
    if path and int(os.environ.get("RANK", "0")) == 0:
        with open(path, "a") as f:
            f.write(json.dumps({k: (float(v) if hasattr(v, "__float__") else str(v)) for k, v in data.items()}) + "\n")


def finish(*args, **kwargs):
    return None
