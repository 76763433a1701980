"""The on-disk hand-off on the shipped datasets/lrs3 sample rows (fixtures copied to
tests/golden/lrs3_handoff): manifest parsing, unit dictionary, the trimming rule and the wav writer."""
import io
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(__file__), "golden", "lrs3_handoff")
# dataset_multi_input.py:219-241 applied to the five test rows (SURVEY.md 8c "Fixtures", probed on the reference)
EXPECT = {"test/UmvOgW6iV2s/00007": (214, 428, 68480), "test/UmvOgW6iV2s/00001": (124, 248, 39680),
          "test/UmvOgW6iV2s/00002": (63, 126, 20160), "test/UmvOgW6iV2s/00004": (178, 356, 56960),
          "test/62cNtvx6P8E/00001": (76, 152, 24320)}


def test_manifest_and_trimming_rule(pkg):
    ho = pkg.hand_off
    root_line, rows = ho.parse_manifest(os.path.join(FIX, "label", "test.tsv"))
    assert root_line.endswith("datasets/lrs3")          # the author's absolute path: callers override it
    assert [r.uid for r in rows] == list(EXPECT)
    assert rows[0].n_video == 107 and rows[0].n_audio == 68608
    code_dict = ho.load_code_dict(os.path.join(FIX, "label", "dict.unt.txt"))
    assert len(code_dict) == 200 and code_dict["17"] == 17
    for r in rows:
        feats, cut = ho.load_item(FIX, r, code_dict)
        u, t, n = EXPECT[r.uid]
        assert feats["code"].shape == (u,) and feats["code"].dtype == np.int64
        assert feats["mel"].shape == (80, t) and feats["mel"].flags["C_CONTIGUOUS"]
        assert feats["spkr"].shape == (256,) and feats["spkr"].dtype == np.float32
        assert cut == n and t == 2 * u
        assert 0 <= feats["code"].min() and feats["code"].max() < 200
    # the first row is the cfg1 golden input
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg1.npz"))
    feats, _ = ho.load_item(FIX, rows[0], code_dict)
    assert np.array_equal(feats["code"], z["code"]) and np.array_equal(feats["mel"], z["mel"])
    assert np.array_equal(feats["spkr"], z["spkr"])
    assert ho.output_name(rows[0]) == "pred_wav/UmvOgW6iV2s/00007"


def test_manifest_filters_and_alignment_check(pkg, tmp_path):
    ho = pkg.hand_off
    _, rows = ho.parse_manifest(os.path.join(FIX, "label", "test.tsv"), max_keep=89, min_keep=37)
    assert [r.n_video for r in rows] == [62, 89, 37]
    bad = tmp_path / "bad.tsv"
    bad.write_text("/root\nx\tv.mp4\taudio/x.wav\t10\t6400\n")
    (tmp_path / "bad.unt").write_text("1 2 3\n")
    with pytest.raises(ValueError):
        ho.parse_manifest(str(bad))
    assert ho.trim_lengths(68608, 214, 429) == (214, 428, 68480)
    assert ho.trim_lengths(1000, 10, 10) == (3, 6, 960)


def test_wav_writer_matches_scipy(pkg, tmp_path):
    from scipy.io import wavfile
    x = (np.random.default_rng(0).standard_normal(1234) * 8000).astype(np.int16)
    p = tmp_path / "a" / "x.wav"
    pkg.hand_off.write_wav_int16(str(p), x)
    buf = io.BytesIO()
    wavfile.write(buf, 16000, x)
    assert p.read_bytes() == buf.getvalue()
    rate, y = wavfile.read(str(p))
    assert rate == 16000 and np.array_equal(x, y)


def test_stage1_mel_deinterleave_matches_reference_expression(pkg):
    """model.py:209-212 of the stage-1 model: (B, T, 160) -> (B, 2T, 80); frame 2t is the even-indexed bins."""
    import torch
    x = torch.arange(2 * 3 * 160, dtype=torch.float32).reshape(2, 3, 160)
    y = pkg.hand_off.stage1_mel_to_frames(x)
    assert y.shape == (2, 6, 80)
    assert torch.equal(y[:, 0::2], x[:, :, 0::2]) and torch.equal(y[:, 1::2], x[:, :, 1::2])
