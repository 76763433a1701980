"""One tap-offset convolution case run through l2s_debug_conv and checked against
torch's conv1d / conv_transpose1d on the same (bf16-rounded) operands.  Used by
tests/test_gpu_conv.py and tools/conv_probe.py (each tcgen05 variant runs in its
own process there: a faulting kernel poisons the CUDA context)."""
import ctypes as C
import json
import sys

import torch
import torch.nn.functional as F


def run_case(pkg, impl, cin, cout, k, dil=1, lin=300, batch=2, up=0, act_bf16=True, use_res=True, use_acc=True,
             div=3.0, slope=0.1, seed=0, knobs=None):
    """Returns dict(max_err_raw, max_err_act, ref_max, tol).  up > 0 makes it a
    ConvTranspose1d(cin, cout, k, stride=up, padding=(k-up)//2) in polyphase form."""
    cabi = pkg._cabi
    lib = cabi.load()
    for kk, vv in (knobs or {}).items():
        assert lib.l2s_debug_set(kk.encode(), int(vv)) == 0, kk
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(seed)
    adt = torch.bfloat16 if act_bf16 else torch.float32
    cin_pad = (cin + 15) // 16 * 16 if cin <= 64 else (cin + 63) // 64 * 64
    x = torch.randn(batch, lin, cin, generator=g).to(adt)          # channels-last
    if up:
        p = (k - up) // 2
        w = (torch.randn(cin, cout, k, generator=g) / (cin * k / up) ** 0.5).to(adt)
        ntaps, ntot, mrows = (k + up - 1) // up, up * cout, lin + 1
        tap_off = [-m for m in range(ntaps)]
        packed = torch.zeros(ntaps, ntot, cin_pad, dtype=adt)
        for m in range(ntaps):
            for r in range(up):
                j = r + m * up
                if j < k:
                    packed[m, r * cout:(r + 1) * cout, :cin] = w[:, :, j].t()
        bias = torch.randn(cout, generator=g) * 0.1
        bias_p = bias.repeat(up)
        lout = lin * up
        out_shift, out_valid = -p * cout, lout * cout
        ref = F.conv_transpose1d(x.float().transpose(1, 2).to(dev), w.float().to(dev), bias.to(dev), stride=up, padding=p)
    else:
        w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(adt)
        ntaps, ntot, mrows = k, cout, lin
        tap_off = [(j - (k - 1) // 2) * dil for j in range(k)]
        packed = torch.zeros(ntaps, ntot, cin_pad, dtype=adt)
        packed[:, :, :cin] = w.permute(2, 0, 1)
        bias = torch.randn(cout, generator=g) * 0.1
        bias_p = bias
        lout = lin
        out_shift, out_valid = 0, lout * cout
        ref = F.conv1d(x.float().transpose(1, 2).to(dev), w.float().to(dev), bias.to(dev), dilation=dil,
                       padding=dil * (k - 1) // 2)
    ref = ref.transpose(1, 2).contiguous()                          # (B, lout, cout)
    res = torch.randn(batch, lout, cout, generator=g) if use_res else None
    acc = torch.randn(batch, lout, cout, generator=g) if use_acc else None
    if res is not None:
        ref = ref + res.to(dev)
    if acc is not None:
        ref = ref + acc.to(dev)
    ref = ref / div
    ref_act = F.leaky_relu(ref, slope)

    xp = torch.zeros(batch, lin, cin_pad, dtype=adt)
    xp[:, :, :cin] = x
    d_x, d_w, d_b = xp.to(dev), packed.to(dev).contiguous(), bias_p.float().to(dev)
    d_res = res.to(dev) if res is not None else None
    d_acc = acc.to(dev) if acc is not None else None
    out_raw = torch.full((batch, lout, cout), float("nan"), device=dev)
    out_act = torch.full((batch, lout, cout), float("nan"), device=dev).to(adt)
    d = cabi.ConvDesc()
    d.inp, d.w, d.bias = d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr()
    d.out_raw, d.out_act = out_raw.data_ptr(), out_act.data_ptr()
    d.res = d_res.data_ptr() if d_res is not None else None
    d.acc_in = d_acc.data_ptr() if d_acc is not None else None
    d.act_bf16 = 1 if act_bf16 else 0
    d.batch, d.lin, d.cin_pad, d.ntaps, d.ntot, d.mrows = batch, lin, cin_pad, ntaps, ntot, mrows
    for i, t in enumerate(tap_off):
        d.tap_off[i] = t
    d.out_shift, d.out_valid, d.scale, d.slope = out_shift, out_valid, div, slope
    err = C.create_string_buffer(512)
    torch.cuda.synchronize()
    st = lib.l2s_debug_conv(C.byref(d), impl, 0, None, err, 512)
    if st != 0:
        raise RuntimeError(f"l2s_debug_conv status {st}: {err.value.decode()}")
    torch.cuda.synchronize()
    e_raw = float((out_raw - ref).abs().max())
    e_act = float((out_act.float() - ref_act).abs().max())
    ref_max = float(ref.abs().max())
    # fp32 accumulation of exactly representable products: only summation order differs
    tol_raw = 2e-4 * max(1.0, ref_max)
    tol_act = (1.0 / 128 if act_bf16 else 2e-4) * max(1.0, ref_max)
    return dict(max_err_raw=e_raw, max_err_act=e_act, ref_max=ref_max, tol_raw=tol_raw, tol_act=tol_act,
                ok=bool(e_raw == e_raw and e_raw <= tol_raw and e_act <= tol_act))


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    # argv[1]: JSON list of [name, kwargs]; one RESULT line per case, in order.  A
    # faulting case ends the process (the context is gone), later cases stay unreported.
    pkg_ = ge.load_package()
    for name, kw in json.loads(sys.argv[1]):
        try:
            res = run_case(pkg_, **kw)
        except Exception as exc:     # noqa: BLE001
            print("RESULT " + json.dumps([name, {"ok": False, "error": repr(exc)[-300:]}]), flush=True)
            break
        print("RESULT " + json.dumps([name, res]), flush=True)
