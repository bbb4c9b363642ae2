"""Summarise a trace from tools/trace_pair.py. usage: python tools/trace_show.py gpurun_out/trace_128.json [bn]"""
import json, sys
d = json.load(open(sys.argv[1]))
n, H, cin, cout = d["shape"]
tr = d["trace"]
bn = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if cout >= 256 else cout)
steps = 3 * (cin // 64)
t0 = min(v for r in tr for v in r if v > 0)
print("kernel us", d["us"], "steps/item", steps)
mma = [v - t0 for v in tr[0] if v > 0]
per_item = 2 + 3 * steps
items = len(mma) // per_item
print("MMA warp: items", items)
tot_acc = tot_a = tot_issue = 0
for it in range(items):
    s = mma[it * per_item:(it + 1) * per_item]
    acc_wait = s[1] - s[0]
    a_wait = sum(s[2 + 3 * k + 1] - s[2 + 3 * k] for k in range(steps))
    issue = sum(s[2 + 3 * k + 2] - s[2 + 3 * k + 1] for k in range(steps))
    tot_acc += acc_wait; tot_a += a_wait; tot_issue += issue
    if it < 6 or it >= items - 2:
        print(f"  item {it}: start {s[0]} acc_empty wait {acc_wait} a_full waits {a_wait} issue(+b waits) {issue} total {s[-1]-s[0]}",
              "a waits:", [s[2 + 3 * k + 1] - s[2 + 3 * k] for k in range(steps)][:12])
print(f"  totals: acc_empty {tot_acc} a_full {tot_a} issue {tot_issue} span {mma[-1]-mma[0]}")
for role, name in ((3, "epi0"), (4, "epi1")):
    e = [v - t0 for v in tr[role] if v > 0]
    k = len(e) // 4
    w = sum(e[4 * i + 1] - e[4 * i] for i in range(k))
    rd = sum(e[4 * i + 2] - e[4 * i + 1] for i in range(k))
    rest = sum(e[4 * i + 3] - e[4 * i + 2] for i in range(k))
    print(f"{name}: items {k} acc_full wait {w} (avg {w // max(k,1)}) tmem-drain {rd} (avg {rd // max(k,1)}) after-release {rest} (avg {rest // max(k,1)}) end {e[-1] if e else 0}")
for role, name in ((1, "prodA"), (2, "prodB")):
    e = [v - t0 for v in tr[role] if v > 0]
    k = len(e) // 2
    w = sum(e[2 * i + 1] - e[2 * i] for i in range(k))
    print(f"{name}: loads {k} empty-wait total {w} span {e[-1]-e[0] if e else 0}; first 12 waits {[e[2*i+1]-e[2*i] for i in range(min(k,12))]}")
