"""TEST INFRASTRUCTURE: run one of the reference's scripts with the reference's OWN modules (no substitution), on the
CPU or whatever torch picks — the only help it gets is an in-memory stand-in for the missing fvcore package.
    python tests/ref_runner.py REF_ROOT SCRIPT [script args...]"""
import os
import runpy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if __name__ == "__main__":
    ref_root, script, rest = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3:]
    from multimodal_siamese_cd_b200.config import install_fvcore_stub
    install_fvcore_stub()
    sys.path.insert(0, ref_root)
    os.chdir(ref_root)
    sys.argv = [script, *rest]
    runpy.run_path(os.path.join(ref_root, script), run_name="__main__")
