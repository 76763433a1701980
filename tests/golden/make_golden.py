"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference classes (MelCodeGenerator, CodeGenerator) from
/root/reference with a stub for the one missing third-party import
(matplotlib, pulled in by speech-resynthesis/utils.py:13-17), loads the
state dict made by oracle.vocoder_oracle.init_state_dict with strict=True,
calls .eval().remove_weight_norm() exactly like inference_server.py:120-123,
and stores fp64 forward outputs (rounded to fp32 for storage) plus a few
intermediate taps.  Nothing here is read at test time except the .npz files.
"""
import hashlib
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import vocoder_oracle as vo  # noqa: E402


def import_reference():
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", m)
    sys.modules.setdefault("matplotlib.pylab", types.ModuleType("matplotlib.pylab"))
    sys.path.insert(0, os.path.join(REF, "multi_input_vocoder"))
    from models_multi_input import MelCodeGenerator  # noqa
    sys.path.insert(0, os.path.join(REF, "speech-resynthesis"))
    from models import CodeGenerator  # noqa
    from utils import AttrDict  # noqa
    return MelCodeGenerator, CodeGenerator, AttrDict


def sd_digest(sd):
    hsh = hashlib.sha256()
    for k in sorted(sd):
        hsh.update(k.encode())
        hsh.update(sd[k].contiguous().numpy().tobytes())
    return hsh.hexdigest()


def load_cfg1_inputs():
    """First row of datasets/lrs3/label/test.tsv with the trimming rule of
    dataset_multi_input.py:219-241 restated (the dataset module itself needs
    librosa/soundfile, absent here)."""
    ds = os.path.join(REF, "datasets/lrs3")
    with open(os.path.join(ds, "label/test.tsv")) as f:
        rows = f.read().strip().split("\n")[1:]
    with open(os.path.join(ds, "label/test.unt")) as f:
        unts = f.read().strip().split("\n")
    uid, _, _, _, nsamp = rows[0].split("\t")
    nsamp = int(nsamp)
    units = np.array([int(t) for t in unts[0].split("|")[-1].split()], dtype=np.int64)
    mel = np.load(os.path.join(ds, "mel", uid + ".npy"))
    spk = np.load(os.path.join(ds, "spk_emb", uid + ".npy"))
    u = min(nsamp // 320, len(units))
    t = min(nsamp // 160, len(mel))
    cut = min(160 * t, 320 * u)
    mel = mel[: cut // 160]
    units = units[: cut // 320]
    return uid, units, mel.T.copy(), spk


def run_reference(cls, AttrDict, hdict, sd, fp64=True, taps=None, **inputs):
    h = AttrDict(dict(hdict))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = cls(h)
        g.load_state_dict(sd, strict=True)
        g.eval()
        g.remove_weight_norm()
    if fp64:
        g = g.double()
        inputs = {k: (v.double() if v.is_floating_point() else v) for k, v in inputs.items()}
    hooks = []
    if taps is not None:
        def mk(name):
            def hook(mod, inp, out):
                taps[name] = out.detach().clone()
            return hook
        hooks.append(g.conv_pre.register_forward_hook(mk("conv_pre")))
        for i, up in enumerate(g.ups):
            hooks.append(up.register_forward_hook(mk(f"ups.{i}")))
        for n, rb in enumerate(g.resblocks):
            hooks.append(rb.register_forward_hook(mk(f"resblocks.{n}")))
    with torch.no_grad():
        y = g(**inputs)
    for hk in hooks:
        hk.remove()
    return y


def main():
    MelCodeGenerator, CodeGenerator, AttrDict = import_reference()
    meta = {"torch": torch.__version__, "cases": {}}
    h = vo.shipped_config()

    sds = {st: vo.init_state_dict(h, seed=1234, style=st) for st in ("ref", "trained")}
    meta["state_dict_sha256"] = {st: sd_digest(sd) for st, sd in sds.items()}

    # ---- cfg1: the shipped sample utterance -------------------------------
    uid, units, mel, spk = load_cfg1_inputs()
    code_t = torch.from_numpy(units).unsqueeze(0)
    mel_t = torch.from_numpy(mel).unsqueeze(0)
    spk_t = torch.from_numpy(spk).unsqueeze(0)
    out = {"code": units, "mel": mel.astype(np.float32), "spkr": spk.astype(np.float32)}
    for st, sd in sds.items():
        y = run_reference(MelCodeGenerator, AttrDict, h, sd, code=code_t, mel=mel_t, spkr=spk_t)
        out[f"wave_{st}"] = y.squeeze().numpy().astype(np.float32)
        y32 = run_reference(MelCodeGenerator, AttrDict, h, sd, fp64=False, code=code_t, mel=mel_t, spkr=spk_t)
        meta["cases"][f"cfg1_{st}"] = {
            "uid": uid, "T": int(mel.shape[1]), "U": int(len(units)),
            "fp32_vs_fp64_max_abs": vo.max_abs(y, y32), "fp32_vs_fp64_snr_db": vo.snr_db(y, y32),
            "rms": float(y.pow(2).mean().sqrt()), "max": float(y.abs().max()),
        }
    np.savez_compressed(os.path.join(HERE, "cfg1.npz"), **out)

    # ---- small synthetic batch, both styles, with taps ---------------------
    code, melb, spkb = vo.synthetic_inputs(2, 16, seed=52)
    out = {"code": code.numpy(), "mel": melb.numpy(), "spkr": spkb.numpy()}
    for st, sd in sds.items():
        taps = {}
        y = run_reference(MelCodeGenerator, AttrDict, h, sd, taps=taps, code=code, mel=melb, spkr=spkb)
        out[f"wave_{st}"] = y.numpy().astype(np.float32)
        if st == "trained":
            for name in ("conv_pre", "ups.0", "resblocks.0", "resblocks.2", "ups.1", "ups.4", "resblocks.14"):
                out[f"tap_{name}"] = taps[name].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "small_b2_t16.npz"), **out)

    # ---- ragged-ish edge sizes: shortest legal (T=2) and an odd tile tail ---
    for frames in (2, 6):
        code, melb, spkb = vo.synthetic_inputs(1, frames, seed=7 + frames)
        y = run_reference(MelCodeGenerator, AttrDict, h, sds["trained"], code=code, mel=melb, spkr=spkb)
        np.savez_compressed(os.path.join(HERE, f"edge_t{frames}.npz"), code=code.numpy(), mel=melb.numpy(),
                            spkr=spkb.numpy(), wave_trained=y.numpy().astype(np.float32))

    # ---- unit-only parent CodeGenerator ------------------------
    hu = vo.unit_only_config()
    sdu = vo.init_state_dict(hu, seed=1234, style="trained", unit_only=True)
    meta["state_dict_sha256"]["unit_only_trained"] = sd_digest(sdu)
    gen = torch.Generator().manual_seed(52)
    code = torch.randint(0, 200, (2, 12), generator=gen)
    spk_id = torch.randint(0, 200, (2, 1), generator=gen)
    y = run_reference(CodeGenerator, AttrDict, hu, sdu, code=code, spkr=spk_id)
    np.savez_compressed(os.path.join(HERE, "unit_only_b2_u12.npz"), code=code.numpy(), spkr=spk_id.numpy(),
                        wave_trained=y.numpy().astype(np.float32))

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
