import numpy as np


def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_):
    """HTK-style triangular mel filterbank (n_mels, 1 + n_fft // 2).  Only feeds the dataset's loss mel, which
    inference discards (dataset_multi_input.py:275-277)."""
    fmax = fmax or sr / 2.0
    hz2mel = lambda f: 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)
    pts = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    fb = np.zeros((n_mels, freqs.size), dtype=np.float32)
    for i in range(n_mels):
        lo, mid, hi = pts[i], pts[i + 1], pts[i + 2]
        up = (freqs - lo) / max(mid - lo, 1e-9)
        down = (hi - freqs) / max(hi - mid, 1e-9)
        fb[i] = np.maximum(0.0, np.minimum(up, down))
    return fb
