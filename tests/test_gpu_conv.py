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


@pytest.mark.parametrize("shape", SHAPES)
def test_tcgen05_conv(shape):
    # own process: a faulting kernel would poison this process's CUDA context
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "convcase.py"), json.dumps(dict(shape, impl=1))],
                       capture_output=True, text=True, timeout=300)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, (p.stdout + p.stderr)[-800:]
    r = json.loads(line[-1][7:])
    assert r["ok"], r
