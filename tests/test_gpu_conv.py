"""The tap-offset convolution building block (csrc/conv_common.cuh) through
l2s_debug_conv: CUDA-core kernel in both operand types, and the tcgen05 kernel,
against torch's conv1d / conv_transpose1d on identical operands."""
import json
import os
import subprocess
import sys

import pytest

from convcase import run_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = [
    dict(cin=64, cout=64, k=3, dil=1, lin=300),
    dict(cin=128, cout=128, k=11, dil=5, lin=700),
    dict(cin=256, cout=256, k=7, dil=3, lin=260),
    dict(cin=32, cout=32, k=7, dil=3, lin=3000),
    dict(cin=16, cout=16, k=11, dil=5, lin=5000),
    dict(cin=336, cout=512, k=7, dil=1, lin=400, use_res=False, use_acc=False, div=1.0),
    dict(cin=512, cout=256, k=11, up=5, lin=200, use_res=False, use_acc=False, div=1.0),
    dict(cin=256, cout=128, k=8, up=4, lin=300, use_res=False, use_acc=False, div=1.0),
    dict(cin=32, cout=16, k=4, up=2, lin=1500, use_res=False, use_acc=False, div=1.0),
    dict(cin=16, cout=16, k=3, dil=1, lin=5, batch=1),
]


@pytest.mark.parametrize("act_bf16", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_simt_conv(pkg, shape, act_bf16):
    r = run_case(pkg, impl=0, act_bf16=act_bf16, **shape)
    assert r["ok"], r


def test_tcgen05_conv_tf32(pkg):
    """fp32 operands through the same tcgen05 kernel as kind::tf32 (10-bit mantissa products)."""
    for shape in (SHAPES[0], SHAPES[3], SHAPES[4], SHAPES[7]):
        r = run_case(pkg, impl=1, act_bf16=False, **shape)
        # operands are truncated to tf32 inside the tensor core: ~2^-10 relative per product
        assert r["max_err_raw"] <= 2e-2 * max(1.0, r["ref_max"]), r
        assert r["max_err_act"] <= 2e-2 * max(1.0, r["ref_max"]), r


@pytest.fixture(scope="module")
def tc_results():
    # own process: a faulting kernel would poison this process's CUDA context
    batch = [[str(i), dict(sh, impl=1)] for i, sh in enumerate(SHAPES)]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "convcase.py"), json.dumps(batch)],
                       capture_output=True, text=True, timeout=600)
    out = {}
    for line in p.stdout.splitlines():
        if line.startswith("RESULT "):
            name, res = json.loads(line[7:])
            out[name] = res
    out["_tail"] = (p.stdout + p.stderr)[-800:]
    return out


@pytest.mark.parametrize("idx", range(len(SHAPES)))
def test_tcgen05_conv(tc_results, idx):
    assert str(idx) in tc_results, tc_results["_tail"]
    assert tc_results[str(idx)]["ok"], tc_results[str(idx)]
