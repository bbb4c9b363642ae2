"""SASS-mnemonic census of the tensor-core kernels in libb200cd.so (no GPU needed: cuobjdump disassembles the cubin).
Writes profiles/r02_sass_mnemonics.txt: per kernel instantiation the counts of the instructions that prove the sm_100a
path — UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), UTMALDG / UTMASTG (TMA loads / stores), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit, .MULTICAST across the CTA pair), SYNCS (mbarrier), plus HMMA / legacy tensor ops (must be 0).
usage: python tools/sass_count.py"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "multimodal_siamese_cd_b200" / "libb200cd.so"
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "HMMA", "IMMA",
            "WGMMA", "REDUX", "SHFL", "ATOM", "RED"]


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    kernels[cur][p] += 1
                    break
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = [f"# cuobjdump -sass {LIB.name}: instruction counts per kernel (sm_100a)", "# columns: " + " ".join(PATTERNS) + " total"]
    for (name, c), dm in zip(kernels.items(), demangled):
        short = re.sub(r"\(anonymous namespace\)::", "", dm)
        short = re.sub(r"^void b200cd::", "", short).split("(")[0]
        if not any(c[p] for p in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "IMMA", "WGMMA")):
            continue
        lines.append(f"{short:70s} " + " ".join(f"{p}={c[p]}" for p in PATTERNS if c[p]) + f" total={c['_total']}")
    lines.append(f"# kernels in the library: {len(kernels)}; legacy tensor instructions (HMMA/IMMA/WGMMA) anywhere: "
                 f"{sum(c['HMMA'] + c['IMMA'] + c['WGMMA'] for c in kernels.values())}")
    (ROOT / "profiles" / "r02_sass_mnemonics.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:12]), "\n...", lines[-1])


if __name__ == "__main__":
    sys.exit(main())
