"""GPU probe of the split-bf16 ("precise") mode: per-kernel checks, then end-to-end parity against the exact fp32 and
fp64 oracles next to the fast mode. Writes gpurun_out/precise_probe.json. Usage: python tools/precise_probe.py [--e2e-only]"""
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

out = {"kernels": {}, "e2e": {}}
if "--e2e-only" not in sys.argv:
    import gpu_checks_hp as gch  # noqa: E402
    for name, fn in gch.HP_CHECKS.items():
        try:
            r = fn()
        except Exception as e:  # noqa: BLE001
            r = {"ok": False, "error": f"{type(e).__name__}: {e}", "tb": traceback.format_exc()[-600:]}
        out["kernels"][name] = r
        print(name, "OK" if r.get("ok") else "FAIL", {k: v for k, v in r.items() if k not in ("ok", "tb")}, flush=True)

import e2e_checks as E  # noqa: E402

SMALL, FULL = (64, 128), (64, 128, 256, 512)
cases = {
    "siamese_small": dict(mtype="siameseunet", cin=4, topo=SMALL, B=3, H=32, W=32, kind="supervised"),
    "dtsiamese_small_fused": dict(mtype="dtsiameseunet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="dualtask", path="fused", steps=3),
    "whatevernet_small": dict(mtype="whatevernet", cin=6, topo=SMALL, B=3, H=32, W=32, kind="mmcr"),
    "siamese_full_256": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", steps=2),
    "siamese_full_256_corr": dict(mtype="siameseunet", cin=4, topo=FULL, B=2, H=256, W=256, kind="supervised", path="fused", corr=True),
    "dualstream_full_128": dict(mtype="dualstreamunet", cin=6, topo=FULL, B=2, H=128, W=128, kind="supervised", path="fused"),
}
if "--quick" in sys.argv:
    cases = {k: cases[k] for k in ("siamese_small", "siamese_full_256")}
for name, kw in cases.items():
    for prec in ("precise", "fast"):
        t0 = time.time()
        try:
            r = E.run_case(**kw, precision=prec, fp64=True, skip_q=True)
        except Exception as e:  # noqa: BLE001
            r = {"error": f"{type(e).__name__}: {e}", "tb": traceback.format_exc()[-1500:]}
        r["seconds"] = time.time() - t0
        out["e2e"][f"{name}:{prec}"] = r
        keep = ("logits_x", "logits_d", "loss_x", "grads_x", "grads_d", "floor_logits", "floor_grads", "mask_flips_x",
                "margin3_flips_x", "floor_mask_flips", "error", "tb", "seconds")
        print(name, prec, json.dumps({k: r[k] for k in keep if k in r}, default=str), flush=True)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "precise_probe.json").write_text(json.dumps(out, indent=1, default=str))
