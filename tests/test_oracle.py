"""The CPU oracle (oracle/vocoder_oracle.py) against golden vectors that were
produced by the unmodified reference classes (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import vocoder_oracle as vo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _digest(sd):
    hsh = hashlib.sha256()
    for k in sorted(sd):
        hsh.update(k.encode())
        hsh.update(sd[k].contiguous().numpy().tobytes())
    return hsh.hexdigest()


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def weights():
    h = vo.shipped_config()
    out = {}
    for st in ("ref", "trained"):
        sd = vo.init_state_dict(h, seed=1234, style=st)
        out[st] = (sd, vo.fold_weight_norm(sd))
    return h, out


def test_state_dict_is_reproducible(meta, weights):
    _, w = weights
    for st in ("ref", "trained"):
        assert _digest(w[st][0]) == meta["state_dict_sha256"][st]


def test_state_dict_key_set(weights):
    # 298 weight-normed tensors, 201 after folding
    _, w = weights
    sd, folded = w["ref"]
    assert len(sd) == 298
    assert len(folded) == 201
    assert sd["ups.0.weight_g"].shape == (512, 1, 1)      # ConvTranspose1d: per input channel
    assert sd["conv_pre.weight_g"].shape == (512, 1, 1)   # Conv1d: per output channel
    assert sd["resblocks.14.convs2.2.weight_v"].shape == (16, 16, 11)


@pytest.mark.parametrize("style", ["ref", "trained"])
def test_cfg1_matches_reference(weights, style):
    h, w = weights
    z = np.load(os.path.join(GOLDEN, "cfg1.npz"))
    code = torch.from_numpy(z["code"]).unsqueeze(0)
    mel = torch.from_numpy(z["mel"]).unsqueeze(0)
    spk = torch.from_numpy(z["spkr"]).unsqueeze(0)
    assert code.shape == (1, 214) and mel.shape == (1, 80, 428)
    y = vo.mel_code_generator_forward(w[style][1], h, code, mel, spk, dtype=torch.float64)
    assert y.shape == (1, 1, 68480)
    ref = torch.from_numpy(z[f"wave_{style}"]).double()
    # golden is fp64 rounded to fp32 for storage
    assert vo.max_abs(ref, y) < 2e-7
    assert vo.snr_db(ref, y) > 120.0


def test_small_batch_and_taps(weights):
    h, w = weights
    z = np.load(os.path.join(GOLDEN, "small_b2_t16.npz"))
    code, mel, spk = (torch.from_numpy(z[k]) for k in ("code", "mel", "spkr"))
    # the synthetic input generator is itself deterministic
    c2, m2, s2 = vo.synthetic_inputs(2, 16, seed=52)
    assert torch.equal(code, c2) and torch.equal(mel, m2) and torch.equal(spk, s2)
    for st in ("ref", "trained"):
        taps = {}
        y = vo.mel_code_generator_forward(w[st][1], h, code, mel, spk, dtype=torch.float64, taps=taps)
        assert vo.max_abs(torch.from_numpy(z[f"wave_{st}"]), y) < 2e-7
        if st == "trained":
            for k in z.files:
                if k.startswith("tap_"):
                    assert vo.max_abs(torch.from_numpy(z[k]), taps[k[4:]]) < 5e-6, k


@pytest.mark.parametrize("frames", [2, 6])
def test_edge_lengths(weights, frames):
    h, w = weights
    z = np.load(os.path.join(GOLDEN, f"edge_t{frames}.npz"))
    y = vo.mel_code_generator_forward(w["trained"][1], h, torch.from_numpy(z["code"]),
                                      torch.from_numpy(z["mel"]), torch.from_numpy(z["spkr"]))
    assert y.shape == (1, 1, 160 * frames)
    assert vo.max_abs(torch.from_numpy(z["wave_trained"]), y) < 2e-7


def test_unit_only_variant(meta):
    hu = vo.unit_only_config()
    sd = vo.init_state_dict(hu, seed=1234, style="trained", unit_only=True)
    assert _digest(sd) == meta["state_dict_sha256"]["unit_only_trained"]
    z = np.load(os.path.join(GOLDEN, "unit_only_b2_u12.npz"))
    y = vo.code_generator_forward(vo.fold_weight_norm(sd), hu, torch.from_numpy(z["code"]),
                                  torch.from_numpy(z["spkr"]))
    assert y.shape == (2, 1, 320 * 12)
    assert vo.max_abs(torch.from_numpy(z["wave_trained"]), y) < 2e-7


def test_error_behaviour_matches_reference(weights):
    h, w = weights
    code, mel, spk = vo.synthetic_inputs(1, 8)
    with pytest.raises(RuntimeError):   # torch.cat size mismatch, models_multi_input.py:73
        vo.mel_code_generator_forward(w["ref"][1], h, code[:, :3], mel, spk)
    with pytest.raises(IndexError):     # embedding index out of range
        vo.mel_code_generator_forward(w["ref"][1], h, code + 200, mel, spk)


def test_flop_model_matches_baseline():
    # 247.746 MFLOP per mel frame, 16.083 GFLOP per audio-s unit-only
    assert abs(vo.algorithmic_flops_per_frame(vo.shipped_config()) / 1e6 - 247.746) < 0.01
    per_unit = vo.algorithmic_flops_per_frame(vo.unit_only_config(), unit_only=True)
    assert abs(per_unit * 50 / 1e9 - 16.0828) < 0.001
