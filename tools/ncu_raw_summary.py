"""Per-kernel summary of an `ncu --page raw --csv` export: launches, total time, DRAM bytes per kernel name, and the
slowest launches. usage: python tools/ncu_raw_summary.py raw.csv [top_n]"""
import csv
import re
import sys

MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
        "ns": 1e-3, "us": 1.0, "ms": 1e3}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ik, idur = names.index("Kernel Name"), names.index("gpu__time_duration.sum")
    ir = names.index("dram__bytes_read.sum") if "dram__bytes_read.sum" in names else None
    iw = names.index("dram__bytes_write.sum") if "dram__bytes_write.sum" in names else None
    ig = names.index("Grid Size")
    out = []
    for r in rows[hdr + 2:]:
        if len(r) <= idur:
            continue
        m = re.search(r"(\w+)(<[^(]*>)?\(", r[ik])
        short = (m.group(1) + (m.group(2) or "")) if m else r[ik][:60]
        us = float(r[idur].replace(",", "")) * MULT.get(units[idur], 1.0)
        by = 0.0
        if ir is not None:
            by = float(r[ir].replace(",", "")) * MULT.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * MULT.get(units[iw], 1.0)
        out.append({"id": int(r[0]), "name": short, "us": us, "dram": by, "grid": r[ig], "full": r[ik]})
    return out


if __name__ == "__main__":
    L = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    mine = [x for x in L if "b200cd" in x["full"]]
    tot = {}
    for x in mine:
        t = tot.setdefault(x["name"], [0, 0.0, 0.0])
        t[0] += 1
        t[1] += x["us"]
        t[2] += x["dram"]
    print(f"{'kernel':60s} {'n':>4s} {'us':>9s} {'MB dram':>9s} {'GB/s':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {v[0]:4d} {v[1]:9.1f} {v[2] / 1e6:9.1f} {v[2] / v[1] / 1e3 if v[1] else 0:7.0f}")
    print(f"{'sum':60s} {len(mine):4d} {sum(v[1] for v in tot.values()):9.1f}")
    print("\nslowest launches")
    for x in sorted(mine, key=lambda x: -x["us"])[:top]:
        print(f"  #{x['id']:4d} {x['name']:56s} {x['us']:8.1f} us {x['dram'] / 1e6:8.1f} MB grid {x['grid']}")
