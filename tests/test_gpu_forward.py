"""Parity of the CUDA path (through the C ABI, via the drop-in classes) against
the CPU oracle and the golden vectors made from the unmodified reference.

Tolerances (stated per mode, vs the fp64 reference run; floors sit within 3 dB of what is measured on B200,
profiles/r02_parity.txt; 1 int16 LSB of the waveform callers quantise = 3.05e-5):
  fp32 mode  CUDA-core FFMA, fp32 storage      : SNR >= 113 dB, max-abs <= 2e-5 (< 1 LSB)   measured 116.5-132.8 dB
  tf32 mode  tcgen05 kind::tf32, fp32 storage  : SNR >= 58 dB,  max-abs <= 4e-3 (131 LSB)   measured 59.8-63.1 dB
  bf16 mode  tcgen05 bf16 operands, fp32 accum,
             fp32 residual stream              : SNR >= 43 dB, max-abs <= 2e-2 (655 LSB) on trained-like weights,
                                                 measured 45.8-47.0 dB / <= 1.2e-2 (393 LSB)
  unit-only variant (shorter stack, rates [5,4,4,2,2]): bf16 >= 39 dB (measured 42.0), tf32 >= 57 dB (measured 59.8)
Unit-table indexing is bit-exact in every mode.
"""
import os

import numpy as np
import pytest
import torch

from oracle import vocoder_oracle as vo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"

TOL = {"fp32": dict(snr=113.0, max_abs=2e-5), "bf16": dict(snr=43.0, max_abs=2e-2), "tf32": dict(snr=58.0, max_abs=4e-3)}
TOL_UNIT_ONLY = {"fp32": dict(snr=113.0, max_abs=2e-5), "bf16": dict(snr=39.0, max_abs=2e-2), "tf32": dict(snr=57.0, max_abs=4e-3)}


def make_gen(pkg, h, sd, precision, cls="MelCodeGenerator", fold=True):
    g = getattr(pkg, cls)(pkg.AttrDict(h))
    g.load_state_dict(sd, strict=True)
    g.eval()
    if fold:
        g.remove_weight_norm()
    g.set_precision(precision)
    return g.to(DEV)


@pytest.fixture(scope="module")
def weights():
    h = vo.shipped_config()
    return h, {st: vo.init_state_dict(h, seed=1234, style=st) for st in ("ref", "trained")}


def check(ref, y, precision, what, tol=None):
    tol = tol or TOL
    y = y.detach().cpu()
    assert y.shape == ref.shape, what
    assert torch.isfinite(y).all(), what
    snr, ma = vo.snr_db(ref, y), vo.max_abs(ref, y)
    print(f"[parity] {what} {precision}: snr {snr:.2f} dB max-abs {ma:.3e} = {ma * 32768.0:.1f} int16 LSB")
    assert snr >= tol[precision]["snr"], f"{what}: SNR {snr:.1f} dB"
    assert ma <= tol[precision]["max_abs"], f"{what}: max-abs {ma:.3e}"
    return snr, ma


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
@pytest.mark.parametrize("style", ["trained", "ref"])
def test_cfg1_golden(pkg, weights, precision, style):
    """configs[0]: the shipped datasets/lrs3 sample utterance (T=428) against the
    fp64 output of the reference classes."""
    h, sds = weights
    z = np.load(os.path.join(GOLDEN, "cfg1.npz"))
    g = make_gen(pkg, h, sds[style], precision)
    y = g(code=torch.from_numpy(z["code"]).unsqueeze(0).to(DEV), mel=torch.from_numpy(z["mel"]).unsqueeze(0).to(DEV),
          spkr=torch.from_numpy(z["spkr"]).unsqueeze(0).to(DEV))
    ref = torch.from_numpy(z["wave_" + style]).view(1, 1, -1)
    if style == "ref" and precision in ("bf16", "tf32"):
        # bias-dominated output (rms 0.1, nearly constant): SNR is not informative, bound the error
        assert vo.max_abs(ref, y.cpu()) <= 2e-3
    else:
        check(ref, y, precision, f"cfg1/{style}")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_small_batch_taps(pkg, weights, precision):
    """B=2, T=16 golden with per-layer taps: conv_pre, ups.0, first MRF stage, final wave."""
    h, sds = weights
    z = np.load(os.path.join(GOLDEN, "small_b2_t16.npz"))
    g = make_gen(pkg, h, sds["trained"], precision)
    lib = pkg._cabi.load()
    code, mel, spkr = (torch.from_numpy(z[k]).to(DEV) for k in ("code", "mel", "spkr"))
    tol = 1e-4 if precision == "fp32" else 6e-2
    try:
        lib.l2s_debug_set(b"stop_after_stage", 0)
        g(code=code, mel=mel, spkr=spkr)
        pre = g.debug_tap("conv_pre_act", (2, 16, 512), DEV)
        ref_pre = torch.nn.functional.leaky_relu(torch.from_numpy(z["tap_conv_pre"]), 0.1).transpose(1, 2)
        assert float((pre - ref_pre).abs().max()) <= tol * max(1.0, float(ref_pre.abs().max()))
        ups0 = g.debug_tap("ups", (2, 80, 256), DEV)
        ref_u = torch.from_numpy(z["tap_ups.0"]).transpose(1, 2)
        assert float((ups0 - ref_u).abs().max()) <= tol * max(1.0, float(ref_u.abs().max()))
    finally:
        lib.l2s_debug_set(b"stop_after_stage", -1)
    y = g(code=code, mel=mel, spkr=spkr)
    check(torch.from_numpy(z["wave_trained"]), y, precision, "small_b2_t16")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
def test_edge_shortest_utterance(pkg, weights, precision):
    """U=1, T=2: every layer is shorter than its receptive field (all padding)."""
    h, sds = weights
    z = np.load(os.path.join(GOLDEN, "edge_t2.npz"))
    g = make_gen(pkg, h, sds["trained"], precision)
    y = g(code=torch.from_numpy(z["code"]).to(DEV), mel=torch.from_numpy(z["mel"]).to(DEV),
          spkr=torch.from_numpy(z["spkr"]).to(DEV))
    check(torch.from_numpy(z["wave_trained"]), y, precision, "edge_t2")
    z6 = np.load(os.path.join(GOLDEN, "edge_t6.npz"))
    y = g(code=torch.from_numpy(z6["code"]).to(DEV), mel=torch.from_numpy(z6["mel"]).to(DEV),
          spkr=torch.from_numpy(z6["spkr"]).to(DEV))
    check(torch.from_numpy(z6["wave_trained"]), y, precision, "edge_t6")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
def test_unit_only_variant(pkg, precision):
    """Parent CodeGenerator.forward (rates [5,4,4,2,2], speaker id table)."""
    h = vo.unit_only_config()
    sd = vo.init_state_dict(h, seed=1234, style="trained", unit_only=True)
    z = np.load(os.path.join(GOLDEN, "unit_only_b2_u12.npz"))
    g = make_gen(pkg, h, sd, precision, cls="CodeGenerator")
    y = g(code=torch.from_numpy(z["code"]).to(DEV), spkr=torch.from_numpy(z["spkr"]).to(DEV))
    check(torch.from_numpy(z["wave_trained"]), y, precision, "unit_only", tol=TOL_UNIT_ONLY)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
def test_oracle_seeded_batch(pkg, weights, precision):
    """Seeded synthetic batch (BASELINE.md section 4 distribution) vs the CPU oracle in fp64."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(3, 100, seed=52)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr, dtype=torch.float64)
    g = make_gen(pkg, h, sds["trained"], precision)
    y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    check(ref, y, precision, "synthetic_b3_t100")


def test_lazy_fold_matches_removed_weight_norm(pkg, weights):
    """forward before remove_weight_norm() folds g*v/||v|| on the fly (train.py-style use)."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(1, 20, seed=3)
    a = make_gen(pkg, h, sds["trained"], "fp32", fold=True)(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    b = make_gen(pkg, h, sds["trained"], "fp32", fold=False)(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    # the lazy fold runs where the parameters live (cuda), remove_weight_norm() folded on the CPU here
    assert float((a - b).abs().max()) <= 2e-6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unit_gather_bit_exact(pkg, weights, precision):
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(2, 64, seed=11)
    code[0, 0], code[1, -1] = 0, 199
    g = make_gen(pkg, h, sds["trained"], precision)
    lib = pkg._cabi.load()
    try:
        lib.l2s_debug_set(b"embed_tap", 1)
        g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
        emb = g.debug_tap("embed", (2, 32, 128), DEV)
    finally:
        lib.l2s_debug_set(b"embed_tap", 0)
    assert torch.equal(emb, sds["trained"]["dict.weight"][code])


def test_mel_fp16_input_promotes(pkg, weights):
    """pred_mel may arrive as float16: same result as the fp32 copy of those values."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(1, 40, seed=5)
    mel16 = mel.half()
    g = make_gen(pkg, h, sds["trained"], "fp32")
    a = g(code=code.to(DEV), mel=mel16.to(DEV), spkr=spkr.to(DEV))
    b = g(code=code.to(DEV), mel=mel16.float().to(DEV), spkr=spkr.to(DEV))
    assert torch.equal(a, b)


def test_int16_output(pkg, weights):
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(2, 40, seed=7)
    g = make_gen(pkg, h, sds["trained"], "fp32")
    y, y16 = g.forward_int16(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    expect = (y.squeeze(1) * 32768.0).clamp(-32768, 32767).cpu().numpy().astype("int16")   # inference.py:79-81
    assert np.array_equal(y16.cpu().numpy(), expect)


def test_error_conventions(pkg, weights):
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "fp32")
    code, mel, spkr = vo.synthetic_inputs(1, 20, seed=1)
    with pytest.raises(RuntimeError):           # torch.cat size mismatch, models_multi_input.py:73
        g(code=code.to(DEV), mel=mel[:, :, :18].contiguous().to(DEV), spkr=spkr.to(DEV))
    with pytest.raises(KeyError):
        g(code=code.to(DEV), spkr=spkr.to(DEV))
    with pytest.raises(RuntimeError):           # no CPU fallback
        g(code=code, mel=mel, spkr=spkr)
    g.strict_index_check = True
    bad = code.clone()
    bad[0, 3] = 200
    with pytest.raises(IndexError):
        g(code=bad.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))   # flag was cleared


def test_out_of_range_ids_are_reported_by_default(pkg, weights):
    """Default (non-strict) mode: an out-of-range unit id is never silent (IndexError / device assert).  The
    forward that saw it does not synchronise; the error surfaces at check_index_errors(), or -- late -- at the next
    forward on the same device, which then does no work.  After that the generator is usable again."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "fp32")
    code, mel, spkr = vo.synthetic_inputs(1, 20, seed=1)
    bad = code.clone()
    bad[0, 5] = -7
    assert not g.strict_index_check
    y = g(code=bad.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))          # clamped, flagged, no sync
    assert torch.isfinite(y).all()
    with pytest.raises(IndexError):
        g.check_index_errors(DEV)
    g.check_index_errors(DEV)                                              # cleared by the report
    g(code=bad.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    torch.cuda.synchronize()
    with pytest.raises(IndexError):                                        # reported late by the next forward
        g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    want = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr)
    check(want, g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)), "fp32", "forward after a reported index error")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_upsampler_kernel_wider_than_three_strides(pkg, precision):
    """Edge config: ConvTranspose1d with k > 3u (u = 2, k = 8, padding 3 > u): the last output positions
    come from polyphase row lin + 1, which the launch must cover.  Against the oracle."""
    h = vo.shipped_config(upsample_rates=[2, 4, 2], upsample_kernel_sizes=[8, 16, 4], upsample_initial_channel=128,
                          resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]])
    sd = vo.init_state_dict(h, seed=7, style="trained")
    code, mel, spkr = vo.synthetic_inputs(2, 26, seed=17)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sd), h, code, mel, spkr)
    g = make_gen(pkg, h, sd, precision)
    y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    assert y.shape == (2, 1, 16 * 26)
    check(ref, y, precision, "k > 3u upsamplers")


def test_concurrent_forwards_from_two_threads(pkg, weights):
    """inference_server.py runs Flask's threaded dev server: /vocoder handlers may call the shared generator from any
    thread.  Two threads hammering one generator (each on its own stream) must get exactly the single-threaded bits."""
    import threading
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "bf16")
    jobs = [vo.synthetic_inputs(2, 40 + 8 * (i % 3), seed=200 + i) for i in range(8)]
    want = [g(code=c.to(DEV), mel=m.to(DEV), spkr=s.to(DEV)).cpu() for c, m, s in jobs]
    got, errs = [None] * len(jobs), []

    def worker(ids):
        try:
            st = torch.cuda.Stream(DEV)
            with torch.cuda.stream(st):
                for _ in range(3):
                    for i in ids:
                        c, m, s = jobs[i]
                        got[i] = g(code=c.to(DEV), mel=m.to(DEV), spkr=s.to(DEV)).cpu()
        except Exception as e:                      # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(range(k, len(jobs), 2),)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for a, b in zip(want, got):
        assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch_invariance(pkg, weights, precision):
    """Utterances are independent: a batched forward equals the per-utterance forwards bit for bit."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(4, 60, seed=9)
    g = make_gen(pkg, h, sds["trained"], precision)
    y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    for i in range(4):
        yi = g(code=code[i:i + 1].to(DEV), mel=mel[i:i + 1].to(DEV), spkr=spkr[i:i + 1].to(DEV))
        assert torch.equal(yi[0], y[i])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_chunked_long_form_equals_full(pkg, weights, precision):
    """Config 4 property: chunks with a 24-frame halo reproduce the unchunked forward
    (interior exactly; allow last-bit differences from tile-boundary summation order)."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(1, 700, seed=13)
    g = make_gen(pkg, h, sds["trained"], precision)
    full = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    chunked = pkg.vocode_long(g, code.to(DEV), mel.to(DEV), spkr.to(DEV), core=200)
    assert chunked.shape == full.shape
    assert float((chunked - full).abs().max()) <= (1e-6 if precision == "fp32" else 1e-6)


@pytest.mark.parametrize("frames", [6, 100, 428])
def test_fused_resblock_steps_equal_unfused(pkg, weights, frames):
    """The fused (c1 -> smem -> c2) kernel must give the SAME bits as the two separate tcgen05 convs:
    identical operands, identical per-row accumulation order."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(2, frames, seed=21)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    try:
        lib.l2s_debug_set(b"fuse_branch", 0)      # the whole-ResBlock kernels round differently: tested below
        lib.l2s_debug_set(b"fuse_pairs", 1)
        a = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        assert g.launch_count(2, frames, DEV) == 53
        lib.l2s_debug_set(b"fuse_pairs", 0)
        b = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        assert g.launch_count(2, frames, DEV) == 98
    finally:
        lib.l2s_debug_set(b"fuse_pairs", 1)
        lib.l2s_debug_set(b"fuse_branch", 1)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("frames", [6, 100, 428])
def test_whole_resblock_kernels_match_steps(pkg, weights, frames):
    """Stages with C <= 64 run one kernel per ResBlock (residual stream in TMEM, c2 accumulating onto it).  The sum
    x + conv is then rounded inside the tensor core instead of after it, so bits differ from the step-by-step
    kernels; what must hold: the MRF output of every stage agrees to bf16-noise level (tolerances below: 60 / 50 /
    45 dB after the three narrow stages, measured 76 / 59 / 52), the first two stages are untouched (bit equal), and
    the waveform keeps the bf16 tolerance against the oracle."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(2, frames, seed=21)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    chans = [256, 128, 64, 32, 16]
    rates = [5, 20, 40, 80, 160]
    floors = [None, None, 60.0, 50.0, 45.0]
    try:
        for stage in range(5):
            taps = []
            for fb in (0, 1):
                lib.l2s_debug_set(b"fuse_branch", fb)
                lib.l2s_debug_set(b"stop_after_stage", stage)
                g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
                taps.append(g.debug_tap("mrf", (2, frames * rates[stage], chans[stage]), device=DEV))
            assert torch.isfinite(taps[1]).all()
            if floors[stage] is None:
                assert torch.equal(taps[0], taps[1])
            else:
                assert vo.snr_db(taps[0], taps[1]) >= floors[stage], (stage, vo.snr_db(taps[0], taps[1]))
        lib.l2s_debug_set(b"stop_after_stage", -1)
        lib.l2s_debug_set(b"fuse_branch", 1)
        y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).cpu()
        assert g.launch_count(2, frames, DEV) == 33
    finally:
        lib.l2s_debug_set(b"stop_after_stage", -1)
        lib.l2s_debug_set(b"fuse_branch", 1)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr)
    check(ref, y, "bf16", f"whole-ResBlock kernels, {frames} frames")


@pytest.mark.parametrize("knob,value", [("res_mode", 1), ("res_mode", 2), ("res_quad_pct", 200), ("res_quad_pct", 0), ("res_cg2", 0), ("res_cg2", 1), ("res_msub", 2),
                                        ("res_msub", 4), ("res_wide", 0), ("narrow_par", 1), ("res_iss2", 1)])
def test_whole_resblock_tilings_agree(pkg, weights, knob, value):
    """Tile size, CTAs per SM (1 / 2 / 4) and epilogue warps per CTA (8 / 4) of the whole-ResBlock kernel change the
    schedule, not the arithmetic of any output element: the waveform must not change by a bit.  narrow_par: branches 2 and 1
    of a narrow stage on two streams, branch 0 last with a two-input epilogue, ((x_0 + x_1) + x_2) / 3 like the serial chain."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(3, 150, seed=33)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    try:
        lib.l2s_debug_set(b"pack", 0)             # the tap-by-tap whole-ResBlock kernel (the time-packed one is tested below)
        a = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        lib.l2s_debug_set(knob.encode(), value)
        b = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
    finally:
        lib.l2s_debug_set(b"pack", 1)
        lib.l2s_debug_set(b"res_mode", 0)
        lib.l2s_debug_set(b"res_msub", 8)
        lib.l2s_debug_set(b"res_quad_pct", 115)
        lib.l2s_debug_set(b"res_cg2", 4)
        lib.l2s_debug_set(b"res_wide", 1)
        lib.l2s_debug_set(b"narrow_par", 0)
        lib.l2s_debug_set(b"res_iss2", 0)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("extra", [{}, {"res_cg2": 0}, {"res_wide": 0}, {"res_ng": 4}, {"res_ng": 1}, {"res_msub": 2}, {"res_tb": 1, "res_gmax": 2},
                                   {"res_skew_iss2": 0}])
@pytest.mark.parametrize("batch,frames", [(3, 150), (1, 34)])
def test_skewed_resblock_schedule_is_bit_identical(pkg, weights, extra, batch, frames):
    """resq_tc.cuh (two S slabs, per-granule barriers, head / tail weight-stage groups walked granule by granule, the next
    tile's slab loaded under the last conv): a different SCHEDULE of the same MMAs and epilogue arithmetic as res_tc_kernel
    -- every accumulator still receives its taps in ascending order -- so the waveform must not change by a bit, with CTA
    pairs or without, 8 or 16 epilogue warps, 1 / 2 / 4 granules, small tiles, one-tap weight stages."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=33)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    defaults = {"pack": 1, "res_mode": 0, "res_skew": 0, "res_cg2": 4, "res_wide": 1, "res_ng": 2, "res_msub": 8, "res_tb": 0, "res_gmax": 0,
                "res_skew_iss2": 1}
    try:
        lib.l2s_debug_set(b"pack", 0)             # tap-by-tap whole-ResBlock kernels on every narrow stage
        lib.l2s_debug_set(b"res_mode", 2)         # one CTA per SM everywhere: the plans the skewed schedule replaces
        for k, v in extra.items():
            if k not in ("res_ng", "res_tb", "res_gmax", "res_skew_iss2"):
                lib.l2s_debug_set(k.encode(), v)
        a = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        lib.l2s_debug_set(b"res_skew", 1)
        for k, v in extra.items():
            lib.l2s_debug_set(k.encode(), v)
        b = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
    finally:
        for k, v in defaults.items():
            lib.l2s_debug_set(k.encode(), v)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("knob,value", [("pk_mode", 1), ("pk_mode", 2), ("pk_cg2", 0), ("pk_fuse", 0)])
@pytest.mark.parametrize("batch,frames", [(3, 150), (1, 34)])
def test_packed_resblock_variants_agree(pkg, weights, knob, value, batch, frames):
    """Time-packed whole-ResBlock kernel (respk_tc.cuh): tile size / CTAs per SM (pk_mode) and CTA pairs (pk_cg2) change
    where an output element sits in a tile and which phase-major block holds it, not its arithmetic (the same offset
    MMAs in the same order); running the three kernel-size branches of a stage in one launch (pk_fuse) or in three
    changes only where the running branch sum waits: the waveform must not change by a bit."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=33)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    try:
        lib.l2s_debug_set(b"pk_chan", 16 | 32 | 64)     # the time-packed kernel on every narrow stage
        a = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        lib.l2s_debug_set(knob.encode(), value)
        b = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
    finally:
        lib.l2s_debug_set(b"pk_chan", 32)
        lib.l2s_debug_set(b"pk_mode", 0)
        lib.l2s_debug_set(b"pk_cg2", 1)
        lib.l2s_debug_set(b"pk_fuse", 1)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("batch,frames", [(2, 100), (1, 428), (4, 62)])
def test_packed_vs_tap_by_tap_whole_resblock(pkg, weights, batch, frames):
    """The time-packed kernel and the tap-by-tap whole-ResBlock kernel compute the same ResBlocks with a different
    summation order (offset MMAs over P packed time steps vs one MMA per tap) and different bias insertion points:
    MRF outputs of the three narrow stages agree to bf16-noise level, and both keep the bf16 tolerance to the oracle."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=21)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    chans, rates = [256, 128, 64, 32, 16], [5, 20, 40, 80, 160]
    outs = []
    try:
        lib.l2s_debug_set(b"pk_chan", 16 | 32 | 64)
        for stage in (2, 3, 4):
            taps = []
            for pack in (0, 1):
                lib.l2s_debug_set(b"pack", pack)
                lib.l2s_debug_set(b"stop_after_stage", stage)
                g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
                taps.append(g.debug_tap("mrf", (batch, frames * rates[stage], chans[stage]), device=DEV))
            snr = vo.snr_db(taps[0], taps[1])
            print(f"[parity] packed vs tap-by-tap, MRF stage {stage}, {batch}x{frames}: {snr:.1f} dB")
            assert torch.isfinite(taps[1]).all()
            assert snr >= [60.0, 50.0, 45.0][stage - 2], (stage, snr)
        lib.l2s_debug_set(b"stop_after_stage", -1)
        for pack in (0, 1):
            lib.l2s_debug_set(b"pack", pack)
            outs.append(g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).cpu())
    finally:
        lib.l2s_debug_set(b"stop_after_stage", -1)
        lib.l2s_debug_set(b"pack", 1)
        lib.l2s_debug_set(b"pk_chan", 32)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr)
    check(ref, outs[0], "bf16", f"tap-by-tap whole-ResBlock {batch}x{frames}")
    check(ref, outs[1], "bf16", f"time-packed whole-ResBlock {batch}x{frames}")


@pytest.mark.parametrize("knob", ["dual", "cluster", "cg2", "alias_at", "epi_tma", "pdl", "chain"])
def test_fused_kernel_variants_agree(pkg, weights, knob):
    """Two-CTAs-per-SM plans (dual), CTA pairs (cluster) with plain weight multicast or cta_group::2 MMAs (cg2) and
    the other fused-step variants change scheduling only: the waveform must not change by a bit.  chain: steps 1 and 2 of a
    ResBlock launched programmatically, consuming the previous step item by item through completion counters (three
    forwards per setting: eager, graph capture, graph replay)."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(3, 150, seed=33)
    g = make_gen(pkg, h, sds["trained"], "bf16")
    lib = pkg._cabi.load()
    outs = []
    try:
        lib.l2s_debug_set(b"fuse_branch", 0)      # exercise the per-step kernels on every stage
        for v in (0, 1):
            lib.l2s_debug_set(knob.encode(), v)
            for _ in range(3 if knob == "chain" else 1):
                y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
            outs.append(y)
    finally:
        lib.l2s_debug_set(b"chain", 0)
        lib.l2s_debug_set(b"dual", 1)
        lib.l2s_debug_set(b"cluster", 1)
        lib.l2s_debug_set(b"cg2", 1)
        lib.l2s_debug_set(b"alias_at", 1)
        lib.l2s_debug_set(b"epi_tma", 0)
        lib.l2s_debug_set(b"pdl", 0)
        lib.l2s_debug_set(b"fuse_branch", 1)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]), float((outs[0] - outs[1]).abs().max())


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_graph_replay_equals_eager(pkg, weights, precision):
    """The conv chain is run eagerly on the first forward of a shape, captured into a CUDA graph on the second and
    replayed afterwards: all three must give the same bits, also for new inputs of the same shape."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], precision)
    lib = pkg._cabi.load()
    code, mel, spkr = vo.synthetic_inputs(2, 90, seed=41)
    code2, mel2, spkr2 = vo.synthetic_inputs(2, 90, seed=42)
    try:
        lib.l2s_debug_set(b"use_graph", 0)
        ref1 = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone()
        ref2 = g(code=code2.to(DEV), mel=mel2.to(DEV), spkr=spkr2.to(DEV)).clone()
        lib.l2s_debug_set(b"use_graph", 1)
        outs = [g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).clone() for _ in range(3)]
        other = g(code=code2.to(DEV), mel=mel2.to(DEV), spkr=spkr2.to(DEV)).clone()
    finally:
        lib.l2s_debug_set(b"use_graph", 1)
    for o in outs:
        assert torch.equal(o, ref1)
    assert torch.equal(other, ref2)


def test_vocode_manifest_batched_equals_per_utterance(pkg, weights, tmp_path):
    """the batched manifest caller (device-side int16) against the reference's per-utterance flow
    (forward, * 32768, astype int16 on the host, inference.py:79-81) on the shipped sample rows."""
    from scipy.io import wavfile
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "fp32")
    fix = os.path.join(GOLDEN, "lrs3_handoff")
    ho = pkg.hand_off
    paths = ho.vocode_manifest(g, os.path.join(fix, "label", "test.tsv"), str(tmp_path), root=fix, device=DEV)
    _, rows = ho.parse_manifest(os.path.join(fix, "label", "test.tsv"))
    code_dict = ho.load_code_dict(os.path.join(fix, "label", "dict.unt.txt"))
    assert len(paths) == 5
    for r, pth in zip(rows, paths):
        assert pth.endswith(ho.output_name(r) + ".wav")
        feats, n = ho.load_item(fix, r, code_dict)
        y = g(**{k: torch.from_numpy(v).to(DEV).unsqueeze(0) for k, v in feats.items()})
        expect = (y.squeeze() * 32768.0).clamp(-32768, 32767).cpu().numpy().astype("int16")
        rate, got = wavfile.read(pth)
        assert rate == 16000 and got.shape == (n,) and np.array_equal(got, expect)


def test_host_pipeline_equals_direct_forward(pkg, weights):
    """HostPipeline overlaps the copies of neighbouring batches with the forward; every pinned output must equal the
    plain forward of the same batch (different batches, two shapes, more batches than buffer slots)."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "bf16")
    pipe = pkg.HostPipeline(g, DEV)
    batches, outs, want = [], [], []
    for i in range(7):
        frames = 60 if i % 3 else 84
        code, mel, spkr = vo.synthetic_inputs(2, frames, seed=90 + i)
        batches.append((code.pin_memory(), mel.pin_memory(), spkr.pin_memory()))
        outs.append(torch.empty((2, 1, frames * 160), dtype=torch.float32).pin_memory())
    for (c, m, s), o in zip(batches, outs):
        pipe.submit(c, m, s, o)
    pipe.finish()
    for (c, m, s), o in zip(batches, outs):
        assert torch.equal(o, g(code=c.to(DEV), mel=m.to(DEV), spkr=s.to(DEV)).cpu())
    # int16 pinned outputs: the device-side int16 waveform (inference.py:79-81), half the device->host bytes
    outs16 = [torch.empty((2, o.shape[-1]), dtype=torch.int16).pin_memory() for o in outs]
    for (c, m, s), o in zip(batches, outs16):
        pipe.submit(c, m, s, o)
    pipe.finish()
    for o32, o16 in zip(outs, outs16):
        expect = (o32.squeeze(1) * 32768.0).clamp(-32768, 32767).numpy().astype("int16")
        assert np.array_equal(o16.numpy(), expect)


def test_stage1_outputs_straight_into_the_vocoder(pkg, weights):
    """device-resident stage-1 outputs (time-major mel frames, unit ids, speaker embeddings of several
    utterances with different lengths, one mel a frame longer than 2U as the mel head produces) through the batched
    caller must equal the per-utterance forward -> *32768 -> int16 flow bit for bit."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "fp32")
    ho = pkg.hand_off
    mels, units, spks, expect = [], [], [], []
    for i, (frames, extra) in enumerate([(40, 0), (40, 1), (64, 0), (40, 0), (22, 3)]):
        code, mel, spkr = vo.synthetic_inputs(1, frames, seed=80 + i)
        tm = mel[0].transpose(0, 1).contiguous()                        # (T, 80) as stage 1 keeps it
        if extra:
            tm = torch.cat([tm, tm[-extra:]], dim=0)
        mels.append(tm.to(DEV)); units.append(code[0].to(DEV)); spks.append(spkr[0].to(DEV))
        y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
        expect.append((y.squeeze() * 32768.0).clamp(-32768, 32767).to(torch.int16))
    got = ho.vocode_stage1_outputs(g, mels, units, spks, device=DEV)
    assert len(got) == 5
    for a, b in zip(got, expect):
        assert a.dtype == torch.int16 and torch.equal(a, b)
    # ... and against the ORACLE: the utterance whose stage-1 mel is three frames longer than 2U (lengths reconciled as the
    # dataset does) through the CPU restatement of the reference
    code, mel, spkr = vo.synthetic_inputs(1, 22, seed=84)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr)
    want = (ref.squeeze().double() * 32768.0).clamp(-32768, 32767).to(torch.int16)
    lsb = int((got[4].cpu().int() - want.int()).abs().max())
    print(f"[parity] stage-1 hand-off, 22 frames (+3 surplus mel frames) fp32 vs oracle: max {lsb} int16 LSB")
    assert got[4].shape == want.shape and lsb <= 1


@pytest.mark.parametrize("precision", ["bf16", "tf32", "fp32"])
@pytest.mark.parametrize("batch,frames", [(1, 2), (3, 34), (2, 514)])
def test_guard_bands_stay_intact(pkg, weights, precision, batch, frames):
    """No sanitizer on this pool: call the C ABI with the workspace and the output embedded in larger buffers filled
    with a canary pattern.  Every kernel (TMA zero-fill boxes, halo rows, tile tails, polyphase edges) must leave the
    bytes before and after both buffers untouched, and the result must equal the host class's forward."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], precision)
    code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=61)
    cd, md, sp = code.to(DEV), mel.to(DEV), spkr.to(DEV)
    want = g(code=cd, mel=md, spkr=sp)
    eng = g._engine(torch.device(DEV))
    lib = eng.lib
    need = int(lib.l2s_workspace_bytes(eng.handle, batch, frames))
    guard = 1 << 20
    big = torch.full((need + 2 * guard,), 0xA5, dtype=torch.uint8, device=DEV)
    n_out = batch * eng.hop * frames
    out_big = torch.full((n_out + 2 * 4096,), -1234.5, dtype=torch.float32, device=DEV)
    ws_ptr = big.data_ptr() + guard            # torch allocations are 512-byte aligned, the guard keeps that
    out_ptr = out_big.data_ptr() + 4096 * 4
    st = lib.l2s_forward(eng.handle, torch.cuda.current_stream().cuda_stream, cd.data_ptr(), md.data_ptr(), pkg._cabi.F32,
                         sp.data_ptr(), batch, frames // 2, frames, out_ptr, ws_ptr, need)
    pkg._cabi.raise_for(lib, eng.handle, st)
    torch.cuda.synchronize()
    assert bool((big[:guard] == 0xA5).all()) and bool((big[guard + need:] == 0xA5).all()), "workspace guard band overwritten"
    assert bool((out_big[:4096] == -1234.5).all()) and bool((out_big[4096 + n_out:] == -1234.5).all()), "output guard band overwritten"
    got = out_big[4096:4096 + n_out].view(batch, 1, -1)
    assert torch.equal(got, want)


@pytest.mark.parametrize("batch,frames", [(1, 2), (1, 10), (3, 34), (5, 66), (2, 514), (1, 1300)])
def test_odd_shapes_whole_resblock_vs_steps(pkg, weights, batch, frames):
    """Tile tails, utterances shorter than a tile, several tiles per utterance, batch sizes that do not divide the
    grid: the whole-ResBlock kernels and the step kernels must agree to bf16-noise level and both match the oracle."""
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "bf16")
    code, mel, spkr = vo.synthetic_inputs(batch, frames, seed=71)
    lib = pkg._cabi.load()
    outs = []
    try:
        for fb in (0, 1):
            lib.l2s_debug_set(b"fuse_branch", fb)
            outs.append(g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).cpu())
        lib.l2s_debug_set(b"pk_chan", 16 | 32 | 64)     # and with the time-packed kernel on every narrow stage
        outs.append(g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV)).cpu())
    finally:
        lib.l2s_debug_set(b"fuse_branch", 1)
        lib.l2s_debug_set(b"pk_chan", 32)
    assert torch.isfinite(outs[1]).all() and torch.isfinite(outs[2]).all()
    assert vo.snr_db(outs[0], outs[1]) >= 40.0, vo.snr_db(outs[0], outs[1])
    assert vo.snr_db(outs[0], outs[2]) >= 40.0, vo.snr_db(outs[0], outs[2])
    if frames <= 600:
        ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sds["trained"]), h, code, mel, spkr)
        check(ref, outs[1], "bf16", f"odd shape {batch}x{frames}")


def test_cfg2_shape_bf16_vs_fp32_device_reference(pkg, weights):
    """configs[1] at full size (16 x 4 s): bf16 tensor-core path against the fp32
    CUDA-core mode of the same library, plus finiteness and range."""
    h, sds = weights
    code, mel, spkr = vo.synthetic_inputs(16, 400, seed=52)
    a = make_gen(pkg, h, sds["trained"], "fp32")(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    b = make_gen(pkg, h, sds["trained"], "bf16")(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    assert b.shape == (16, 1, 64000)
    assert torch.isfinite(b).all() and float(b.abs().max()) <= 1.0
    snr = vo.snr_db(a.cpu(), b.cpu())
    print(f"[parity] cfg2 bf16 vs fp32-device: snr {snr:.2f} dB max-abs {vo.max_abs(a.cpu(), b.cpu()):.3e}")
    assert snr >= 43.0


@pytest.mark.parametrize("devices", ["twice_gpu0", "all"])
def test_in_process_multi_gpu_dispatcher_equals_single_gpu(pkg, weights, devices):
    """one process, one engine + HostPipeline + host thread per GPU, utterances dealt
    longest-first, int16 back.  The waveforms must equal the per-utterance single-GPU flow bit for bit.  "twice_gpu0"
    runs two workers on cuda:0 (exercises the threading on a one-GPU box); "all" needs >= 2 GPUs."""
    h, sds = weights
    if devices == "all" and torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    devs = [0, 0] if devices == "twice_gpu0" else list(range(torch.cuda.device_count()))
    g = make_gen(pkg, h, sds["trained"], "bf16")
    feats, want = [], []
    for i, frames in enumerate([60, 84, 60, 30, 84, 60, 84, 30, 60, 12, 84]):
        code, mel, spkr = vo.synthetic_inputs(1, frames, seed=400 + i)
        feats.append({"code": code[0].numpy(), "mel": mel[0].numpy(), "spkr": spkr[0].numpy()})
        y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
        want.append((y.squeeze() * 32768.0).clamp(-32768, 32767).to(torch.int16).cpu().numpy())
    mg = pkg.MultiGpuVocoder(g, devices=devs, max_batch=2)
    try:
        for _ in range(2):                                   # second call reuses engines, pipelines, graphs
            got = mg.vocode(feats)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert a.dtype == np.int16 and np.array_equal(a, b)
        # equal-length utterances already stacked on the host: straight from slices of the caller's pinned tensors
        code, mel, spkr = vo.synthetic_inputs(7, 60, seed=77)
        wav = mg.vocode_batch(code.pin_memory(), mel.pin_memory(), spkr.pin_memory())
        for i in range(7):
            y = g(code=code[i:i + 1].to(DEV), mel=mel[i:i + 1].to(DEV), spkr=spkr[i:i + 1].to(DEV))
            assert torch.equal(wav[i], (y.squeeze() * 32768.0).clamp(-32768, 32767).to(torch.int16).cpu())
    finally:
        mg.close()


def test_serve_vocoder_request_equals_per_utterance_flow(pkg, weights, tmp_path):
    """the batched, I/O-overlapped /vocoder request handler against the reference's per-utterance flow
    (forward, * 32768, astype int16 on the host, scipy wav) on the shipped sample rows: identical files."""
    from scipy.io import wavfile
    h, sds = weights
    g = make_gen(pkg, h, sds["trained"], "fp32")
    fix = os.path.join(GOLDEN, "lrs3_handoff")
    ho = pkg.hand_off
    paths = ho.serve_vocoder_request(g, fix, str(tmp_path), device=DEV, max_batch=4, io_threads=4)
    _, rows = ho.parse_manifest(os.path.join(fix, "label", "test.tsv"))
    code_dict = ho.load_code_dict(os.path.join(fix, "label", "dict.unt.txt"))
    assert len(paths) == 5
    for r, pth in zip(rows, paths):
        feats, n = ho.load_item(fix, r, code_dict)
        y = g(**{k: torch.from_numpy(v).to(DEV).unsqueeze(0) for k, v in feats.items()})
        expect = (y.squeeze() * 32768.0).clamp(-32768, 32767).cpu().numpy().astype("int16")
        ref_path = os.path.join(str(tmp_path), "ref.wav")
        wavfile.write(ref_path, 16000, expect[:n])
        assert open(pth, "rb").read() == open(ref_path, "rb").read()
    # ... and against the ORACLE (not only against the CUDA path itself): the two shortest rows through the CPU restatement
    # of the reference, quantised the way inference.py:79-81 does; the written files may differ by one int16 step at most
    folded = vo.fold_weight_norm(sds["trained"])
    for i in sorted(range(len(rows)), key=lambda k: rows[k].n_audio)[:2]:
        feats, n = ho.load_item(fix, rows[i], code_dict)
        ref = vo.mel_code_generator_forward(folded, h, *(torch.from_numpy(feats[k]).unsqueeze(0) for k in ("code", "mel", "spkr")))
        want = (ref.squeeze().double() * 32768.0).clamp(-32768, 32767).numpy().astype("int16")[:n]
        _, got = wavfile.read(paths[i])
        assert got.shape == want.shape
        lsb = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max())
        print(f"[parity] /vocoder request, row {rows[i].uid} ({n} samples) fp32 vs oracle: max {lsb} int16 LSB")
        assert lsb <= 1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("pk_chan", [0, 16 | 32 | 64])
def test_unusual_resblock_config(pkg, precision, pk_chan):
    """A config far from the shipped one: kernel sizes 5 / 9 (the run-time MMA issue path of the time-packed kernel: its
    immediate-operand path covers k = 3 / 7 / 11), two dilations per ResBlock, even dilations (phase-major layouts with
    d = 2 / 4), different dilation lists per branch (so the branches of a stage cannot share one launch), rates [4,4,2].
    With and without the time-packed kernel, against the oracle."""
    h = vo.shipped_config(upsample_rates=[4, 4, 2], upsample_kernel_sizes=[8, 8, 4], upsample_initial_channel=128,
                          resblock_kernel_sizes=[5, 9], resblock_dilation_sizes=[[1, 2], [2, 4]])
    sd = vo.init_state_dict(h, seed=11, style="trained")
    code, mel, spkr = vo.synthetic_inputs(3, 46, seed=19)
    ref = vo.mel_code_generator_forward(vo.fold_weight_norm(sd), h, code, mel, spkr)
    g = make_gen(pkg, h, sd, precision)
    lib = pkg._cabi.load()
    try:
        lib.l2s_debug_set(b"pk_chan", pk_chan)
        y = g(code=code.to(DEV), mel=mel.to(DEV), spkr=spkr.to(DEV))
    finally:
        lib.l2s_debug_set(b"pk_chan", 32)
    assert y.shape == (3, 1, 32 * 46)
    check(ref, y, precision, f"k = 5 / 9, dilations [1,2] / [2,4], pk_chan {pk_chan}")
