"""DRAM traffic of the CTA-pair convolution kernel, launch by launch, for one training step of a BASELINE config.

  # on the GPU box, under ncu (one eager step on ONE stream, so launch order == call order):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -o gpurun_out/r02_step_all python tools/ncu_step_traffic.py run dualstream
  ncu -i gpurun_out/r02_step_all.ncu-rep --page raw --csv > gpurun_out/r02_step_all_raw.csv
  # anywhere:
  python tools/ncu_step_traffic.py join dualstream gpurun_out/r02_step_all_raw.csv gpurun_out/r02_step_calls_dualstream.json

`run` executes the step and writes (a) the list of launches that use fprop_pair_kernel (family, shape tag, algorithmic
FLOPs and bytes) in call order and (b) per-family totals of every other family; `join` pairs (a) with ncu's per-launch
dram__bytes_read.sum + dram__bytes_write.sum and writes profiles/r02_traffic_<family>_<config>.json (what bench.py
reports as roofline.traffic) plus a per-launch table, and puts the DRAM bytes of the remaining kernels (summed by kernel
name) next to the algorithmic bytes of their family in profiles/r02_traffic_step_<config>.json.
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
PAIR_FAMILIES = ("fprop3x3", "dgrad3x3_bnbwd", "gemm1tap", "convT_dgrad", "convT_dgrad_bnbwd")
# kernel-name prefix -> ops.PROFILE family, for the kernels that are not the CTA-pair convolution
NAME_TO_FAMILY = (("wgrad_reduce", "wgrad_reduce"), ("wgrad_kernel", "wgrad"), ("bn_bwd", "bn_bwd"), ("bn_apply", "bn_apply"),
                  ("bn_stats", "bn_stats"), ("bn_finalize", "bn_stats"), ("pack_input", "pack_input"),
                  ("pack_weights", "pack_weights"), ("colsum", "colsum"), ("stat_rowsum", "colsum"), ("head_fwd", "head_fwd"),
                  ("pj_", "pj"))


def run(cfgname: str) -> None:
    import torch

    import bench
    from multimodal_siamese_cd_b200 import networks, ops
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    mtype, cin, B, kind, alpha, _, _ = bench.CONFIGS[cfgname]
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    net.module.use_cuda_graphs = False
    ts = TrainStep(net.module, B, 256, 256, kind=kind, alpha=alpha, device=dev, dp_group=None)
    ts.eng.branch_streams = ts.eng.wgrad_side = False
    g = torch.Generator(device=dev).manual_seed(7)
    xc = 6 if mtype in bench.TWO_STREAM else cin
    ts.eng.x_t1.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    ts.eng.x_t2.copy_(torch.rand(B, xc, 256, 256, device=dev, generator=g))
    for t in ts.targets.values():
        t.copy_((torch.rand(t.shape, device=dev, generator=g) > 0.9).float())
    ops.PROFILE = []
    ts.run()
    torch.cuda.synchronize()
    calls = [{"family": n, "tag": tag, "flops": fl, "algorithmic_bytes": by} for n, fl, by, _e0, _e1, tag in ops.PROFILE
             if n in PAIR_FAMILIES]
    fams: dict = {}
    for n, fl, by, _e0, _e1, _tag in ops.PROFILE:
        f = fams.setdefault("pj" if n.startswith("pj_") else n, {"calls": 0, "flops": 0.0, "algorithmic_bytes": 0.0})
        f["calls"] += 1
        f["flops"] += fl
        f["algorithmic_bytes"] += by
    ops.PROFILE = None
    out = ROOT / "gpurun_out" / f"r02_step_calls_{cfgname}.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps({"pair_calls": calls, "families": fams}, indent=0))
    print(f"{len(calls)} launches of fprop_pair_kernel in one {cfgname} step -> {out}")


def join(cfgname: str, raw_csv: str, calls_json: str) -> None:
    blob = json.loads(Path(calls_json).read_text())
    calls, fams = blob["pair_calls"], blob["families"]
    rows = list(csv.reader(open(raw_csv)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ik, ir, iw = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
    idur = names.index("gpu__time_duration.sum")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    launches = []
    other: dict = {}
    for r in rows[hdr + 2:]:
        if len(r) > max(ir, iw) and "fprop_pair_kernel" not in r[ik]:
            short = r[ik].split("(")[0].split("::")[-1].split("<")[0]
            fam = next((f for pre, f in NAME_TO_FAMILY if short.startswith(pre)), None)
            if fam is not None:
                o = other.setdefault(fam, {"launches": 0, "dram_bytes": 0.0, "us_under_ncu": 0.0})
                o["launches"] += 1
                o["dram_bytes"] += float(r[ir].replace(",", "")) * mult.get(units[ir], 1.0) + \
                    float(r[iw].replace(",", "")) * mult.get(units[iw], 1.0)
                o["us_under_ncu"] += float(r[idur].replace(",", "")) * mult.get(units[idur], 1.0)
        if len(r) > max(ir, iw) and "fprop_pair_kernel" in r[ik]:
            launches.append({"dram_bytes": float(r[ir].replace(",", "")) * mult.get(units[ir], 1.0) +
                             float(r[iw].replace(",", "")) * mult.get(units[iw], 1.0),
                             "us_under_ncu": float(r[idur].replace(",", "")) * mult.get(units[idur], 1.0),
                             "kernel": r[ik].split("(")[0][-60:]})
    assert len(launches) == len(calls), f"{len(launches)} captured launches vs {len(calls)} calls of one step"
    table = [{**c, **l} for c, l in zip(calls, launches)]
    prof = ROOT / "profiles"
    step = {}
    for fam, o in sorted(other.items()):
        a = fams.get(fam, {}).get("algorithmic_bytes", 0.0)
        step[fam] = {**o, "algorithmic_bytes": a, "dram_over_algorithmic": o["dram_bytes"] / a if a else None}
        print(f"{fam:14s} {o['launches']:4d} launches  {o['dram_bytes'] / 1e6:9.1f} MB dram  {a / 1e6:9.1f} MB algorithmic  "
              f"{o['us_under_ncu']:9.1f} us under ncu")
    (prof / f"r02_traffic_step_{cfgname}.json").write_text(json.dumps(
        {"config": cfgname, "what": "one eager training step on one stream under ncu (--clock-control none): DRAM bytes "
         "(dram__bytes_read.sum + dram__bytes_write.sum) summed per kernel family next to the family's algorithmic bytes "
         "(ops.py _Prof)", "families": step}, indent=1))
    (prof / f"r02_traffic_launches_{cfgname}.json").write_text(json.dumps(table, indent=0))
    for fam in PAIR_FAMILIES:
        sel = [t for t in table if t["family"] == fam]
        if not sel:
            continue
        d = {"config": cfgname, "family": fam, "launches": len(sel),
             "dram_bytes_per_launch": sum(t["dram_bytes"] for t in sel) / len(sel),
             "algorithmic_bytes_per_launch": sum(t["algorithmic_bytes"] for t in sel) / len(sel),
             "dram_over_algorithmic": sum(t["dram_bytes"] for t in sel) / sum(t["algorithmic_bytes"] for t in sel),
             "source": "ncu --set full --clock-control none, one eager step on one stream; per launch: "
                       f"profiles/r02_traffic_launches_{cfgname}.json"}
        (prof / f"r02_traffic_{fam}_{cfgname}.json").write_text(json.dumps(d, indent=1))
        print(fam, d["launches"], f"{d['dram_bytes_per_launch'] / 1e6:.1f} MB dram vs {d['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic "
              f"(x{d['dram_over_algorithmic']:.2f})")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        join(sys.argv[2], sys.argv[3], sys.argv[4])
