"""GPU: every kernel through the C ABI against torch fp32 on identical (bf16-rounded) inputs. Tolerances live next to
each check in tests/gpu_checks.py: bf16-output kernels within one bf16 rounding of the fp32 result (rel L2 < 6e-3,
and identical to the rounded reference on all but ~1e-3 of the elements), fp32-output kernels < 2e-3 ... 1e-5."""
import pytest
import torch

import gpu_checks as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gc.ALL_CHECKS))
def test_kernel(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    res = gc.ALL_CHECKS[name]()
    assert res["ok"], res
    if "ulp_frac" in res:
        assert res["ulp_frac"] < res.get("ulp_tol", 5e-3) and res["rel_l2_rounded"] < 5e-4, res
