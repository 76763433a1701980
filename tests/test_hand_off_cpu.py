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


def _groups_equal(a, b):
    assert len(a) == len(b)
    for (ga, ca, ma, sa, wa), (gb, cb, mb, sb, wb) in zip(a, b):
        assert [it[0] for it in ga] == [it[0] for it in gb] and [it[4] for it in ga] == [it[4] for it in gb]
        assert ca.dtype == cb.dtype and ma.dtype == mb.dtype and sa.dtype == sb.dtype
        assert np.array_equal(ca.numpy(), cb.numpy()) and np.array_equal(ma.numpy(), mb.numpy()) and np.array_equal(sa.numpy(), sb.numpy())
        assert wa.shape == wb.shape and wa.dtype == wb.dtype


def test_native_group_loader_matches_numpy_on_the_lrs3_rows(pkg):
    """l2s_io_read_npy_f32 (include/l2s_hand_off.h) against np.load + the Python collate, on the shipped sample rows:
    one group per row (lengths differ) and all rows in one group (the loader must split it by exact length)."""
    ho = pkg.hand_off
    _, rows = ho.parse_manifest(os.path.join(FIX, "label", "test.tsv"))
    cd = ho.load_code_dict(os.path.join(FIX, "label", "dict.unt.txt"))
    for idxs in ([0], [1], [4], [0, 1, 2, 3, 4]):
        nat = ho._load_group(FIX, rows, idxs, cd, pin=False)
        ref = ho._load_group_python(FIX, rows, idxs, cd, pin=False)
        _groups_equal(sorted(nat, key=lambda g: g[2].shape[1]), sorted(ref, key=lambda g: g[2].shape[1]))
    one = ho._load_group(FIX, rows, [0], cd, pin=False)[0]
    assert tuple(one[2].shape) == (1, 428, 80) and tuple(one[1].shape) == (1, 214) and one[0][0][4] == 68480


def _write_dataset(root, specs, mel_dtype=np.float32, spk=None, fortran=False):
    """specs = [(frames on disk, n_audio)] -> a dataset in the reference's layout under root; returns the mel arrays."""
    os.makedirs(os.path.join(root, "label"))
    with open(os.path.join(root, "label", "dict.unt.txt"), "w") as f:
        f.writelines(f"{i} 1\n" for i in range(200))
    rng = np.random.default_rng(3)
    tsv, unt, mels = [root + "\n"], [], []
    for i, (frames, n_audio) in enumerate(specs):
        rel = f"audio/test/s{i % 2}/{i:03d}.wav"
        mel = rng.standard_normal((frames, 80)).astype(mel_dtype)
        emb = rng.standard_normal(256).astype(np.float32) if spk is None else spk
        for sub, arr in (("mel", np.asfortranarray(mel) if fortran else mel), ("spk_emb", emb)):
            path = os.path.join(root, rel.replace("audio/", sub + "/")[:-4] + ".npy")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.save(path, arr)
        nv = n_audio // 640
        tsv.append(f"test/s{i % 2}/{i:03d}\tv.mp4\t{rel}\t{nv}\t{n_audio}\n")
        unt.append(" ".join(str(int(c)) for c in rng.integers(0, 200, 2 * nv)) + "\n")
        mels.append(mel)
    open(os.path.join(root, "label", "test.tsv"), "w").writelines(tsv)
    open(os.path.join(root, "label", "test.unt"), "w").writelines(unt)
    return mels


def test_native_group_loader_equal_lengths_fp16_and_fallbacks(pkg, tmp_path):
    ho = pkg.hand_off
    # equal lengths: the tensors the native reader filled are the batch itself; longer files are cut at n_audio // 160
    root = str(tmp_path / "a")
    mels = _write_dataset(root, [(400, 64000), (403, 64000), (400, 64000)])
    _, rows = ho.parse_manifest(os.path.join(root, "label", "test.tsv"))
    cd = ho.load_code_dict(os.path.join(root, "label", "dict.unt.txt"))
    nat = ho._load_group(root, rows, [0, 1, 2], cd, pin=False)
    assert len(nat) == 1 and tuple(nat[0][2].shape) == (3, 400, 80)
    assert np.array_equal(nat[0][2].numpy()[1], mels[1][:400])
    _groups_equal(nat, ho._load_group_python(root, rows, [0, 1, 2], cd, pin=False))
    # float16 on disk is widened exactly (including subnormals, zeros, infinities)
    root = str(tmp_path / "b")
    mels = _write_dataset(root, [(200, 32000), (200, 32000)], mel_dtype=np.float16)
    special = np.asarray([0.0, -0.0, 5.96e-8, -6.1e-5, 65504.0, np.inf, -np.inf, 1.0009765625], dtype=np.float16)
    m0 = mels[0].copy(); m0[0, :8] = special
    np.save(os.path.join(root, "mel", "test", "s0", "000.npy"), m0)
    _, rows = ho.parse_manifest(os.path.join(root, "label", "test.tsv"))
    nat = ho._load_group(root, rows, [0, 1], cd, pin=False)
    assert nat[0][2].dtype.is_floating_point and nat[0][2].element_size() == 4
    assert np.array_equal(nat[0][2].numpy()[0], m0.astype(np.float32)) and np.array_equal(nat[0][2].numpy()[1], mels[1].astype(np.float32))
    # Fortran-order files are declined by the native reader and read through numpy: same result
    root = str(tmp_path / "c")
    mels = _write_dataset(root, [(100, 16000), (100, 16000)], fortran=True)
    _, rows = ho.parse_manifest(os.path.join(root, "label", "test.tsv"))
    nat = ho._load_group(root, rows, [0, 1], cd, pin=False)
    assert np.array_equal(nat[0][2].numpy()[0], mels[0]) and np.array_equal(nat[0][2].numpy()[1], mels[1])
    # a speaker embedding of the wrong shape raises the reference's check (helpers.py:194), a missing file raises too
    root = str(tmp_path / "d")
    _write_dataset(root, [(100, 16000)], spk=np.zeros((1, 256), np.float32))
    _, rows = ho.parse_manifest(os.path.join(root, "label", "test.tsv"))
    with pytest.raises(ValueError, match="speaker embedding"):
        ho._load_group(root, rows, [0], cd, pin=False)
    os.remove(os.path.join(root, "mel", "test", "s0", "000.npy"))
    with pytest.raises(OSError):
        ho._load_group(root, rows, [0], cd, pin=False)


def test_native_wav_writer_matches_scipy(pkg, tmp_path):
    """l2s_io_write_wav_i16: n files per call, bytes equal scipy.io.wavfile.write (inference.py:164)."""
    import torch
    from scipy.io import wavfile
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.integers(-32768, 32767, (5, 4000), dtype=np.int16))
    ns = [4000, 1, 0, 3999, 1234]
    paths = [str(tmp_path / f"{i}.wav") for i in range(5)]
    pkg.hand_off._write_group_native(paths, x, ns, native_threads=3)
    for i, (p, n) in enumerate(zip(paths, ns)):
        buf = io.BytesIO()
        wavfile.write(buf, 16000, x[i, :n].numpy())
        assert open(p, "rb").read() == buf.getvalue()
    with pytest.raises(OSError):
        pkg.hand_off._write_group_native([str(tmp_path / "missing_dir" / "x.wav")], x, [10])


def test_native_io_argument_checks(pkg):
    import ctypes as C
    cabi = pkg._cabi
    lib = cabi.load()
    assert lib.l2s_io_read_npy_f32(None, 1, None, 0, None, 80, 0, None, 1, None) == cabi.IO_ERR_ARG
    assert lib.l2s_io_read_npy_f32(None, 0, None, 0, None, 80, 0, None, 1, None) == cabi.IO_OK
    assert lib.l2s_io_write_wav_i16(None, 1, None, 0, None, 16000, 1, None) == cabi.IO_ERR_ARG
    buf = np.zeros(80, np.float32)
    rows, cap, bad = np.zeros(1, np.int32), np.asarray([2], np.int32), C.c_int32(-1)
    ip = C.POINTER(C.c_int32)
    paths = (C.c_char_p * 1)(b"/nonexistent/x.npy")
    # the slot is too small for max_rows * cols
    assert lib.l2s_io_read_npy_f32(paths, 1, buf.ctypes.data, 80, cap.ctypes.data_as(ip), 80, 0, rows.ctypes.data_as(ip), 1, C.byref(bad)) == cabi.IO_ERR_ARG
    cap[0] = 1
    assert lib.l2s_io_read_npy_f32(paths, 1, buf.ctypes.data, 80, cap.ctypes.data_as(ip), 80, 0, rows.ctypes.data_as(ip), 1, C.byref(bad)) == cabi.IO_ERR_OPEN
    assert bad.value == 0


def test_native_units_to_ids_matches_code_to_sequence(pkg):
    """l2s_io_units_to_ids against code_to_sequence (dataset_multi_input.py:128-141, collapse_code = False): the shipped
    lrs3 rows, unknown tokens dropped, blank runs / tabs / trailing newline, the cap, an empty line."""
    import ctypes as C
    ho = pkg.hand_off
    lib = pkg._cabi.load()
    fix = FIX
    _, rows = ho.parse_manifest(os.path.join(fix, "label", "test.tsv"))
    code_dict = ho.load_code_dict(os.path.join(fix, "label", "dict.unt.txt"))
    lines = [r.units for r in rows] + ["7  9\t11 nope 3 \n", "", "   ", "x y z", " ".join(str(i % 200) for i in range(1000))]
    want = [ho.code_to_sequence(ln.split(), code_dict) for ln in lines]
    n = len(lines)
    cap = np.asarray([max(len(w), 1) for w in want], dtype=np.int32)
    cap[-1] = 17                                                    # fewer slots than tokens: the count still reports all of them
    out = np.full((n, int(cap.max())), -1, dtype=np.int64)
    n_out = np.zeros(n, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    st = lib.l2s_io_units_to_ids(ho._c_paths(lines), n, ho._dict_tokens(code_dict), len(code_dict), out.ctypes.data, out.strides[0] // 8,
                                 cap.ctypes.data_as(ip), n_out.ctypes.data_as(ip), 3)
    assert st == pkg._cabi.IO_OK
    for k in range(n):
        assert n_out[k] == len(want[k]), k
        m = min(len(want[k]), int(cap[k]))
        assert out[k, :m].tolist() == want[k][:m], k
        assert (out[k, m:] == -1).all()                             # nothing written past the cap / the count
    assert n_out[len(rows)] == 4 and n_out[len(rows) + 1] == 0 and n_out[len(rows) + 3] == 0
    # argument checks
    assert lib.l2s_io_units_to_ids(None, 1, ho._dict_tokens(code_dict), len(code_dict), out.ctypes.data, 4, cap.ctypes.data_as(ip),
                                   n_out.ctypes.data_as(ip), 1) == pkg._cabi.IO_ERR_ARG

