"""ORACLE — TEST INFRASTRUCTURE ONLY. Stages the UNMODIFIED reference tree (its .py files and configs/*.yaml, ~3 k
lines) from /root/reference into oracle/_ref/reference/ so that it can travel to the GPU box with the repo snapshot
(oracle/_ref/ is git-ignored: nothing of the reference enters the history; it is NOT gpurun-ignored).

Users: bench.py --impl reference (times the reference's own modules on the host cores) and
tests/test_reference_scripts.py (runs the reference's training scripts unchanged through the launcher and plainly).
Called by __graft_entry__.build() when /root/reference exists; a no-op otherwise (the staged copy, if any, is kept).
"""
from __future__ import annotations

import shutil
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref" / "reference"


def stage(force: bool = False) -> Path | None:
    if not SRC.exists():
        return DST if DST.exists() else None
    if DST.exists() and not force:
        src_files = sorted(p.relative_to(SRC) for p in SRC.rglob("*.py")) + sorted(p.relative_to(SRC) for p in SRC.rglob("*.yaml"))
        if all((DST / f).exists() and (DST / f).stat().st_size == (SRC / f).stat().st_size for f in src_files):
            return DST
    if DST.exists():
        shutil.rmtree(DST)
    for pat in ("*.py", "*.yaml"):
        for f in SRC.rglob(pat):
            rel = f.relative_to(SRC)
            if any(part.startswith(".") or part == "__pycache__" for part in rel.parts):
                continue
            (DST / rel).parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(f, DST / rel)
    return DST


def staged_root() -> Path | None:
    """The staged tree, or /root/reference itself when that exists (build container), else None."""
    if SRC.exists():
        return SRC
    return DST if DST.exists() else None


if __name__ == "__main__":
    print(stage(force=True))
