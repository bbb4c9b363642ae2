"""N-tile sweep of the CTA-pair convolution kernel over the launch shapes of the BASELINE configs.

For every distinct (variant, mode, out_mode, n_img, H, W, ka, N, precision) the training plans launch (collected by
running one eager step per config with tuning.LOG on), time the launch with each N tile that divides N (64 / 128 / 256)
and with the library's own rule: CUDA event pair per launch, L2 flushed (256 MB memset) before each, median of
`--reps`. Writes the full table (--out, JSON) and the entries where a forced tile beats the library's rule by more than
`--margin` (--emit, the JSON multimodal_siamese_cd_b200/tuning.py loads through B200CD_TILE_TABLE; paste into
tuning.TABLE to make them the default).

    python tools/tile_sweep.py --out gpurun_out/tile_sweep.json --emit gpurun_out/tuned_tiles.json
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def collect(cfgname: str, precision: str) -> list:
    """Launch keys of one training step of `cfgname`."""
    import torch

    import bench
    from multimodal_siamese_cd_b200 import networks, tuning
    from multimodal_siamese_cd_b200.config import synthetic_cfg
    from multimodal_siamese_cd_b200.step import TrainStep
    mtype, cin, B, kind, alpha, _gf, _yaml = bench.CONFIGS[cfgname]
    dev = torch.device("cuda", 0)
    cfg = synthetic_cfg(mtype, in_channels=cin)
    torch.manual_seed(cfg.SEED)
    net = networks.create_network(cfg).to(dev).train()
    net.module.set_precision(precision)
    tuning.LOG = []
    ts = TrainStep(net.module, B, 256, 256, kind=kind, alpha=alpha, device=dev, dp_group=None)
    ts.run()                      # first run is eager: every launch passes through ops.conv_gemm*
    torch.cuda.synchronize()
    keys = [k for k in tuning.LOG]
    tuning.LOG = None
    del ts, net
    torch.cuda.empty_cache()
    return keys


def time_key(key, reps: int, flush) -> dict:
    import torch

    from multimodal_siamese_cd_b200 import ops
    variant, mode, out_mode, n, H, W, ka, N, prec = key
    dev = "cuda"
    taps = 9 if mode == 0 else (1 if mode == 1 else 4)
    km = 3 if prec else 1

    def act(nn_, h, w, c):
        x = torch.randn(nn_, h, w, c, device=dev) * 0.5
        return ops.split_from_float(x) if prec else x.to(torch.bfloat16)

    A = act(n, 2 * H if mode == 2 else H, 2 * W if mode == 2 else W, ka)
    Bw = (torch.randn(N, taps * ka * km, device=dev) / (taps * ka) ** 0.5).to(torch.bfloat16)
    if out_mode == 1:
        cout = N // 4
        out = act(n, 2 * H, 2 * W, cout)
    else:
        out = act(n, H, W, N)
    bias = torch.zeros(N, device=dev)
    G = 1
    cands = [None] + [b for b in (64, 128, 256) if N % b == 0]
    launches = {}
    keep = []           # buffers of every candidate stay alive until all are timed
    for bn in cands:
        if variant in ("stats", "bnbwd"):
            rows, per_cta = ops.conv_stat_rows(n, H, W, ka, N, G, mode=mode, prec=prec, variant=variant, bn=bn)
            if not per_cta:
                continue
            stats = torch.zeros(G * rows * N * 2, device=dev)
            keep.append(stats)
        if variant == "bnbwd":
            r = act(n, H, W, N)
            sc = torch.rand(G, N, device=dev) + 0.5
            sh = torch.randn(G, N, device=dev) * 0.1
            keep += [r, sc, sh]

            def launch(bn=bn, stats=stats, rows=rows, r=r, sc=sc, sh=sh):
                ops.conv_gemm_bnbwd(mode, A, Bw, out, r, sc, sh, stats.view(G, rows, N, 2), G, bn=bn)
        elif variant == "stats":
            def launch(bn=bn, stats=stats):
                ops.conv_gemm(mode, out_mode, A, Bw, out, bias=bias if out_mode == 0 else None, stats=stats, stat_groups=G,
                              prec=prec, bn=bn)
        elif variant == "affine":
            sc = torch.rand(N, device=dev) + 0.5
            sh = torch.randn(N, device=dev) * 0.1
            keep += [sc, sh]

            def launch(bn=bn, sc=sc, sh=sh):
                ops.conv_gemm_affine(mode, A, Bw, out, bias, sc, sh, True, prec=prec, bn=bn)
        else:
            b1 = torch.zeros(N // 4 if out_mode == 1 else N, device=dev)
            keep.append(b1)

            def launch(bn=bn, b1=b1):
                ops.conv_gemm(mode, out_mode, A, Bw, out, bias=b1, prec=prec, bn=bn)
        launches["auto" if bn is None else str(bn)] = launch
    for f in launches.values():
        for _ in range(3):
            f()
    torch.cuda.synchronize()
    # candidates interleaved rep by rep (clock / thermal drift hits all of them alike), L2 flushed before every launch
    ev = {k: [] for k in launches}
    torch.cuda._sleep(int(6e7))   # ~40 ms spin kernel: the host enqueues every launch below before the first one runs,
    for _ in range(reps):         # so no event interval holds host latency (tensor-map encoding, ctypes)
        for k, f in launches.items():
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            ev[k].append((e0, e1))
    torch.cuda.synchronize()
    ops.device_status()
    res = {}
    for k, pairs in ev.items():
        t = sorted(a.elapsed_time(b) for a, b in pairs)
        res[k] = round(t[len(t) // 2] * 1e3, 2)   # microseconds (median)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="siamese,dualstream,dtsiamese,mmcr")
    ap.add_argument("--precise-configs", default="dualstream")
    ap.add_argument("--reps", type=int, default=21)
    ap.add_argument("--margin", type=float, default=0.03)
    ap.add_argument("--out", default="gpurun_out/tile_sweep.json")
    ap.add_argument("--emit", default="gpurun_out/tuned_tiles.json")
    a = ap.parse_args()
    import torch

    from multimodal_siamese_cd_b200 import tuning
    tuning.ENABLED = False            # collect and time against the library's own rule
    keys, users = {}, {}
    for prec, names in (("fast", a.configs), ("precise", a.precise_configs)):
        for c in [x for x in names.split(",") if x]:
            for k in collect(c, prec):
                keys[k] = keys.get(k, 0) + 1
                users.setdefault(k, set()).add(f"{c}/{prec}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2000):           # ~0.1 s of memsets: clocks settled before the first shape is timed
        flush.zero_()
    torch.cuda.synchronize()
    rows, emit = [], []
    for k in sorted(keys, key=lambda k: (k[8], k[0], k[1], -k[3] * k[4] * k[5], k[6], k[7])):
        t = time_key(k, a.reps, flush)
        torch.cuda.empty_cache()
        variant, mode, out_mode, n, H, W, ka, N, prec = k
        taps = 9 if mode == 0 else (1 if mode == 1 else 4)
        gf = 2.0 * n * H * W * N * taps * ka / 1e9
        forced = {b: v for b, v in t.items() if b != "auto"}
        best = min(forced, key=forced.get) if forced else None
        row = {"key": list(k), "launches_per_step": keys[k], "users": sorted(users[k]), "gflop": round(gf, 2), "us": t,
               "tflops": {b: round(gf / v * 1e3, 1) for b, v in t.items()}, "best": best}
        rows.append(row)
        if best is not None and "auto" in t and forced[best] < t["auto"] * (1.0 - a.margin):
            emit.append(list(k) + [int(best)])
        print(json.dumps(row), flush=True)
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps({"what": __doc__.split("\n")[0], "reps": a.reps, "rows": rows}, indent=1))
    Path(a.emit).write_text(json.dumps(emit))
    saved = sum((r["us"]["auto"] - r["us"][r["best"]]) * r["launches_per_step"] for r in rows
                if r["best"] and "auto" in r["us"] and list(r["key"]) + [int(r["best"])] in emit)
    print(f"{len(emit)} of {len(rows)} shapes re-tiled; isolated time saved over all users' steps: {saved:.0f} us")


if __name__ == "__main__":
    main()
