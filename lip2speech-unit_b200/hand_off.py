"""The on-disk hand-off around the generator: what stage 1 and
``create_dataset.vocoder()`` leave for the vocoder service, and what the service writes back.

    <root>/label/<split>.tsv    first line = dataset root, then  id \\t video \\t audio \\t n_video_frames \\t n_audio_samples
    <root>/label/<split>.unt    one line of space-separated unit tokens per row (an optional ``name|`` prefix is dropped)
    <root>/label/dict.unt.txt   ``token count`` per line; a token's id is its line number
    <root>/mel/<id>.npy         (T', 80) float32 / float16 log-mel, 100 Hz
    <root>/spk_emb/<id>.npy     (256,) float32 speaker embedding
    <out>/pred_wav/<split>/<name>.wav   int16, 16 kHz mono

Mirrors ``multi_input_vocoder/dataset_multi_input.py`` (parse_manifest :40-110, load_code_dict :118-125,
code_to_sequence :128-141, the trimming rule of __getitem__ :219-241, speaker file mapping :174) and the writer
of ``inference.py:152-165`` / ``inference_server.py:133-146`` -- minus what the vocoder never needed at
inference (the ground-truth wav read, the discarded STFT mel, random interval sampling).

``vocode_manifest`` is the batched caller: the reference runs one utterance per forward; here utterances of
equal length are stacked into one forward (the generator has no length masks), and the int16
conversion happens on the device.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

try:
    from . import _cabi
except ImportError:                                       # top-level module use (see models_multi_input.py)
    import _cabi

CODE_HOP, MEL_HOP, SAMPLE_RATE = 320, 160, 16000


@dataclass
class ManifestRow:
    uid: str            # "test/UmvOgW6iV2s/00007"
    audio_rel: str      # "audio/test/UmvOgW6iV2s/00007.wav" (relative to the dataset root)
    n_video: int        # video frames (25 fps): the manifest's size column (items[-2])
    n_audio: int        # audio samples at 16 kHz (items[-1])
    units: str          # unit tokens, space separated


def parse_manifest(manifest_path: str, max_keep: Optional[int] = None, min_keep: Optional[int] = None):
    """Returns (root_line, rows).  Same filtering and the same alignment check (|len(units) - 2 * n_video| <= 2)
    as dataset_multi_input.py:59-77; a misaligned row raises instead of dropping into pdb."""
    code_path = os.path.splitext(manifest_path)[0] + ".unt"
    rows: List[ManifestRow] = []
    with open(manifest_path) as f, open(code_path) as fc:
        root = f.readline().strip()
        for line, line_code in zip(f, fc):
            items = line.strip().split("\t")
            code = line_code.strip().split("|")[-1]
            sz = int(items[-2])
            n_tok = len(code.split())
            if not -2 <= n_tok - sz * 2 <= 2:
                raise ValueError(f"{items[0]}: code length {n_tok} != video length * 2 ({sz * 2})")
            if min_keep is not None and sz < min_keep:
                continue
            if max_keep is not None and sz > max_keep:
                continue
            rows.append(ManifestRow(items[0], items[2], sz, int(items[-1]), code))
    return root, rows


def load_code_dict(path: str) -> Dict[str, int]:
    with open(path) as f:
        codes = [line.rstrip().rsplit(" ", 1)[0] for line in f]
    d = {c: i for i, c in enumerate(codes)}
    if set(d.values()) != set(range(len(d))):
        raise ValueError("duplicate tokens in the unit dictionary")
    return d


def code_to_sequence(tokens: Sequence[str], code_dict: Dict[str, int]) -> List[int]:
    """collapse_code=False branch of dataset_multi_input.py:128-141 (unknown tokens are dropped; the reference warns
    when more than 5 % are)."""
    return [code_dict[c] for c in tokens if c in code_dict]


def trim_lengths(n_audio: int, n_units: int, n_mel: int):
    """The trimming rule of __getitem__ (:219-241): U = min(n // 320, units), T = min(n // 160, mel frames),
    cut = min(160 T, 320 U)  ->  (units kept, mel frames kept, samples)."""
    u = min(n_audio // CODE_HOP, n_units)
    t = min(n_audio // MEL_HOP, n_mel)
    cut = min(t * MEL_HOP, u * CODE_HOP)
    return cut // CODE_HOP, cut // MEL_HOP, cut


def load_item(root: str, row: ManifestRow, code_dict: Dict[str, int], n_audio: Optional[int] = None):
    """code int64 (U,), mel (80, T) in the file's dtype, spkr float32 (256,), and the sample count of the output.
    n_audio defaults to the manifest's sample count (the reference reads the wav for it; they agree)."""
    audio_path = os.path.join(root, row.audio_rel)
    mel = np.load(audio_path.replace("/audio/", "/mel/")[:-4] + ".npy")
    spk = np.load(audio_path.replace("/audio/", "/spk_emb/")[:-4] + ".npy")
    if spk.shape != (256,) or spk.dtype != np.float32:      # helpers.py:194 / create_dataset.py:229
        raise ValueError(f"{row.uid}: speaker embedding must be (256,) float32, got {spk.shape} {spk.dtype}")
    code = np.asarray(code_to_sequence(row.units.split(), code_dict), dtype=np.int64)
    u, t, cut = trim_lengths(row.n_audio if n_audio is None else n_audio, code.shape[0], mel.shape[0])
    return {"code": code[:u], "mel": np.ascontiguousarray(mel[:t].T), "spkr": spk}, cut


def output_name(row: ManifestRow) -> str:
    """pred_wav/<speaker dir>/<file> without extension, as inference.py:156 builds it from the audio path."""
    return os.path.join("pred_wav", *row.audio_rel.split("/")[-2:])[:-4]


def _wav_header(n_samples: int, rate: int = SAMPLE_RATE) -> bytes:
    n = n_samples * 2
    return (b"RIFF" + (36 + n).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little") +
            (1).to_bytes(2, "little") + (1).to_bytes(2, "little") + rate.to_bytes(4, "little") +
            (rate * 2).to_bytes(4, "little") + (2).to_bytes(2, "little") + (16).to_bytes(2, "little") +
            b"data" + n.to_bytes(4, "little"))


def write_wav_int16(path: str, samples: np.ndarray, rate: int = SAMPLE_RATE) -> None:
    """16-bit PCM mono RIFF, what scipy.io.wavfile.write produces for an int16 array (inference.py:164)."""
    samples = np.ascontiguousarray(samples, dtype="<i2")
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "wb") as f:
        f.write(_wav_header(samples.size, rate))
        f.write(samples.tobytes())


@torch.no_grad()
def vocode_batched(generator, feats: Sequence[dict], device="cuda", max_batch: int = 64) -> List[torch.Tensor]:
    """The batched caller.  feats[i] = {"code": (U,) int64, "mel": (80, T = 2U), "spkr": (256,)} as numpy
    arrays or torch tensors (host or device).  The generator has no length masks, so utterances of equal
    length are stacked into one forward (at most max_batch each); host arrays go through one pinned staging buffer per
    batch.  Returns the int16 waveforms (device tensors, 160 T samples each) in input order."""
    def as_tensor(v):
        return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))

    def stack(vals):
        ts = [as_tensor(v) for v in vals]
        if ts[0].is_cuda:
            return torch.stack([t.to(device) for t in ts])
        pinned = torch.empty((len(ts),) + tuple(ts[0].shape), dtype=ts[0].dtype, pin_memory=True)
        torch.stack(ts, out=pinned)
        return pinned.to(device, non_blocking=True)

    by_len: Dict[int, List[int]] = {}
    for i, f in enumerate(feats):
        by_len.setdefault(int(f["mel"].shape[1]), []).append(i)
    out: List[Optional[torch.Tensor]] = [None] * len(feats)
    for _, idxs in sorted(by_len.items()):
        for s in range(0, len(idxs), max_batch):
            grp = idxs[s:s + max_batch]
            _, wav16 = generator.forward_int16(code=stack([feats[i]["code"] for i in grp]),
                                               mel=stack([feats[i]["mel"] for i in grp]),
                                               spkr=stack([feats[i]["spkr"] for i in grp]))
            for j, i in enumerate(grp):
                out[i] = wav16[j]
    return out


@torch.no_grad()
def vocode_manifest(generator, manifest_path: str, out_dir: str, root: Optional[str] = None, device="cuda",
                    max_batch: int = 64, code_dict_path: Optional[str] = None) -> List[str]:
    """Vocode every row of a manifest and write <out_dir>/pred_wav/.../<name>.wav (int16, 16 kHz).
    Rows with the same number of mel frames share one forward; the waveform is quantised on the device."""
    tsv_root, rows = parse_manifest(manifest_path)
    root = root or tsv_root
    code_dict = load_code_dict(code_dict_path or os.path.join(os.path.dirname(manifest_path), "dict.unt.txt"))
    items = [load_item(root, r, code_dict) for r in rows]
    wavs = vocode_batched(generator, [f for f, _ in items], device=device, max_batch=max_batch)
    written = []
    for row, (_, n), w in zip(rows, items, wavs):
        path = os.path.join(out_dir, output_name(row) + ".wav")
        write_wav_int16(path, w[:n].cpu().numpy())
        written.append(path)
    return written


def stage1_mel_to_frames(encoder_out_mel: torch.Tensor) -> torch.Tensor:
    """The mel head of the stage-1 model emits two 80-bin frames per 50 Hz step as (B, T1, 160); the reference
    de-interleaves them to (B, 2 T1, 80) with reshape(B,T,D//2,2).transpose(-1,-2).reshape(B,2T,D//2)
    (multi_target_lip2speech/model.py:209-212).  Same expression, any device."""
    b, t, d = encoder_out_mel.shape
    return encoder_out_mel.reshape(b, t, d // 2, 2).transpose(-1, -2).reshape(b, t * 2, d // 2)


@torch.no_grad()
def vocode_stage1_outputs(generator, mels: Sequence[torch.Tensor], units: Sequence[torch.Tensor],
                          spk_embs: Sequence[torch.Tensor], device="cuda", max_batch: int = 64) -> List[torch.Tensor]:
    """stage-1 outputs straight into the vocoder, no .npy / .unt / HTTP round trip.
    mels[i]: (T_i', 80) time-major frames as the stage-1 generator keeps them (sequence_generator.py:136-139 cuts them
    to 2 * target_length), units[i]: (U_i,) int64 unit ids already mapped through dict.unt.txt, spk_embs[i]: (256,).
    Lengths are reconciled as the dataset does without an audio file: U = min(U_i, T_i' // 2), T = 2 U
    (dataset_multi_input.py:219-241 with n_audio = 320 U).  Tensors may live on the device already."""
    feats = []
    for mel, code, spk in zip(mels, units, spk_embs):
        u = min(int(code.shape[0]), int(mel.shape[0]) // 2)
        feats.append({"code": code[:u].to(torch.int64), "mel": mel[:2 * u].transpose(0, 1).contiguous().float(),
                      "spkr": spk.float()})
    return vocode_batched(generator, feats, device=device, max_batch=max_batch)


def _read_npy(path: str) -> np.ndarray:
    """np.load for the plain little-endian C-order arrays of the hand-off (mel/*.npy, spk_emb/*.npy) without numpy's
    header parser (which dominates for files this small); anything unusual goes through np.load."""
    with open(path, "rb") as f:
        buf = f.read()
    try:
        if buf[:6] != b"\x93NUMPY" or buf[6] != 1:
            raise ValueError
        hl = int.from_bytes(buf[8:10], "little")
        hdr = buf[10:10 + hl].decode("latin1")
        fortran = "'fortran_order': True" in hdr
        if not fortran and "'fortran_order': False" not in hdr:
            raise ValueError
        descr = hdr.split("'descr': '")[1].split("'")[0]
        shape = tuple(int(x) for x in hdr.split("'shape': (")[1].split(")")[0].replace(" ", "").split(",") if x)
        arr = np.frombuffer(buf, dtype=np.dtype(descr), offset=10 + hl)
        return arr.reshape(shape[::-1]).T if fortran else arr.reshape(shape)
    except Exception:
        return np.load(path)


def _plan_groups(rows, max_batch: int, first: int = 8):
    """Row indices grouped by the frame count the manifest implies (n_audio // 160), longest first, <= max_batch each; the
    very first group is small so that the GPU starts while the host is still reading files."""
    order = sorted(range(len(rows)), key=lambda i: (-(rows[i].n_audio // 160), i))
    groups, cur = [], []
    for i in order:
        cap = min(first, max_batch) if not groups else max_batch
        if cur and (len(cur) >= cap or rows[i].n_audio // 160 != rows[cur[0]].n_audio // 160):
            groups.append(cur)
            cur = []
        cur.append(i)
    if cur:
        groups.append(cur)
    return groups


def _stack_group(items, pin: bool):
    """items = [(row index, code, mel (t, bins), spk, cut)] -> one batch per exact length (see _load_group)."""
    by_t: Dict[int, list] = {}
    for it in items:                                  # exact lengths may differ inside a group: split it
        by_t.setdefault(it[2].shape[0], []).append(it)
    out = []
    for t, grp in by_t.items():
        n = len(grp)
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)      # noqa: E731
        code = mk((n, grp[0][1].shape[0]), torch.int64)
        mel = mk((n, t, grp[0][2].shape[1]), torch.float32)      # time-major as on disk; (80, T) (dataset_multi_input.py:243) is made on the device
        spk = mk((n, grp[0][3].shape[0]), torch.float32)
        code_np, mel_np, spk_np = code.numpy(), mel.numpy(), spk.numpy()
        for k, (_, c, m, s_, _) in enumerate(grp):
            code_np[k] = c
            mel_np[k] = m
            spk_np[k] = s_
        wav = mk((n, t * 160), torch.int16)
        out.append((grp, code, mel, spk, wav))
    return out


def _load_group_python(dataset_dir: str, rows, idxs, code_dict, pin: bool = True):
    """_load_group through numpy, one file at a time: the path for anything the native reader declines (Fortran order,
    other dtypes, wrong shapes -- with the reference's own error messages)."""
    items = []
    for i in idxs:
        row = rows[i]
        audio_path = os.path.join(dataset_dir, row.audio_rel)
        mel = _read_npy(audio_path.replace("/audio/", "/mel/")[:-4] + ".npy")
        spk = _read_npy(audio_path.replace("/audio/", "/spk_emb/")[:-4] + ".npy")
        if spk.shape != (256,) or spk.dtype != np.float32:      # helpers.py:194 / create_dataset.py:229
            raise ValueError(f"{row.uid}: speaker embedding must be (256,) float32, got {spk.shape} {spk.dtype}")
        code = np.asarray(code_to_sequence(row.units.split(), code_dict), dtype=np.int64)
        u, t, cut = trim_lengths(row.n_audio, code.shape[0], mel.shape[0])
        items.append((i, code[:u], mel[:t], spk, cut))
    return _stack_group(items, pin)


def _c_paths(paths):
    import ctypes as C
    return (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])


_DICT_TOKENS_CACHE: Dict[int, tuple] = {}


def _dict_tokens(code_dict: Dict[str, int]):
    """The dictionary as a C array of token strings in index order (built once per dict object)."""
    import ctypes as C
    hit = _DICT_TOKENS_CACHE.get(id(code_dict))
    if hit is not None and hit[0] is code_dict:
        return hit[1]
    toks = [b""] * len(code_dict)
    for tok, i in code_dict.items():
        toks[i] = tok.encode()
    arr = (C.c_char_p * len(toks))(*toks)
    if len(_DICT_TOKENS_CACHE) > 8:
        _DICT_TOKENS_CACHE.clear()
    _DICT_TOKENS_CACHE[id(code_dict)] = (code_dict, arr)
    return arr


def _load_group(dataset_dir: str, rows, idxs, code_dict, pin: bool = True, num_mels: int = 80, native_threads: int = 4):
    """Read the .npy files of rows[idxs], apply the trimming rule, and stack rows of equal exact length into (pinned)
    batch tensors: [(items, code (n,U) int64, mel (n,T,80) float32 TIME-MAJOR, spkr (n,256) float32, wav (n,160 T) int16 buffer)]
    with items[k] = (row index, ..., samples to keep).

    The files of the group are read by ONE native call each for mel/ and spk_emb/ (l2s_io_read_npy_f32: native_threads
    host threads, straight into the pinned batch tensors, float16 widened on the fly); a file the native reader declines
    sends the whole group through _load_group_python."""
    import ctypes as C
    lib = _cabi.load()
    n = len(idxs)
    if n == 0:
        return []
    audio = [os.path.join(dataset_dir, rows[i].audio_rel) for i in idxs]
    cap = np.asarray([rows[i].n_audio // MEL_HOP for i in idxs], dtype=np.int32)      # frames the trimming rule can keep at most
    t_max = int(cap.max())
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)      # noqa: E731
    mel = mk((n, max(t_max, 1), num_mels), torch.float32)
    spk = mk((n, 256), torch.float32)
    got = np.zeros(n, dtype=np.int32)
    one = np.ones(n, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    st = lib.l2s_io_read_npy_f32(_c_paths([a.replace("/audio/", "/mel/")[:-4] + ".npy" for a in audio]), n, mel.data_ptr(),
                                 mel.stride(0), cap.ctypes.data_as(ip), num_mels, _cabi.IO_REQUIRE_2D, got.ctypes.data_as(ip),
                                 native_threads, None)
    if st == _cabi.IO_OK:
        got1 = np.zeros(n, dtype=np.int32)
        st = lib.l2s_io_read_npy_f32(_c_paths([a.replace("/audio/", "/spk_emb/")[:-4] + ".npy" for a in audio]), n, spk.data_ptr(),
                                     spk.stride(0), one.ctypes.data_as(ip), 256, _cabi.IO_REQUIRE_1D | _cabi.IO_REQUIRE_F32,
                                     got1.ctypes.data_as(ip), native_threads, None)
    if st != _cabi.IO_OK:
        return _load_group_python(dataset_dir, rows, idxs, code_dict, pin)
    # unit strings -> dictionary indices, one native call for the group (l2s_io_units_to_ids; per row in Python it was a list
    # comprehension over a dict under the GIL: 5-7 ms per 128-utterance request, and the main thread's submit waited for it)
    u_cap = np.asarray([rows[i].n_audio // CODE_HOP for i in idxs], dtype=np.int32)
    u_max = max(int(u_cap.max()), 1)
    ids_all = mk((n, u_max), torch.int64)
    n_ids = np.zeros(n, dtype=np.int32)
    st = lib.l2s_io_units_to_ids(_c_paths([rows[i].units for i in idxs]), n, _dict_tokens(code_dict), len(code_dict), ids_all.data_ptr(),
                                 ids_all.stride(0), u_cap.ctypes.data_as(ip), n_ids.ctypes.data_as(ip), native_threads)
    if st != _cabi.IO_OK:
        return _load_group_python(dataset_dir, rows, idxs, code_dict, pin)
    keep = [trim_lengths(rows[i].n_audio, int(n_ids[k]), int(got[k])) for k, i in enumerate(idxs)]
    if all(kp[:2] == keep[0][:2] for kp in keep) and keep[0][1] == t_max and t_max > 0 and keep[0][0] == u_max:
        # the usual case: every row of the group keeps the same number of frames -- the tensors just read ARE the batch
        wav = mk((n, t_max * MEL_HOP), torch.int16)
        grp = [(i, None, None, None, keep[k][2]) for k, i in enumerate(idxs)]
        return [(grp, ids_all, mel, spk, wav)]
    mel_np, spk_np, ids_np = mel.numpy(), spk.numpy(), ids_all.numpy()
    items = [(i, ids_np[k, :keep[k][0]], mel_np[k, :keep[k][1]], spk_np[k], keep[k][2]) for k, i in enumerate(idxs)]
    return _stack_group(items, pin)


def _write_group_native(paths, wav: torch.Tensor, n_samples, native_threads: int = 4):
    """wav (n, S) int16 host tensor -> n RIFF files (l2s_io_write_wav_i16; bytes as scipy.io.wavfile.write, inference.py:164)."""
    import ctypes as C
    lib = _cabi.load()
    ns = np.asarray(n_samples, dtype=np.int32)
    bad = C.c_int32(-1)
    st = lib.l2s_io_write_wav_i16(_c_paths(paths), len(paths), wav.data_ptr(), wav.stride(0), ns.ctypes.data_as(C.POINTER(C.c_int32)),
                                  SAMPLE_RATE, native_threads, C.byref(bad))
    if st != _cabi.IO_OK:
        raise OSError(f"could not write {paths[bad.value] if 0 <= bad.value < len(paths) else 'wav files'} (status {st})")


@torch.no_grad()
def serve_vocoder_request(generator, dataset_dir: str, out_dir: str, split: str = "test", device="cuda", max_batch: int = 16,
                          io_threads: int = 2, code_dict_path: Optional[str] = None, first_group: int = 8,
                          native_threads: int = 4) -> List[str]:
    """One /vocoder request of the stage-2 service, batched and overlapped.

    The reference handler (multi_input_vocoder/inference_server.py:207-213) re-parses <dataset_dir>/label/<split>.tsv,
    then inference() (:133-146) vocodes ONE item: dataset item -> device -> generator -> * 32768 -> host int16 -> wav
    under <out_dir>/pred_wav/<speaker>/<id>.wav.  Same inputs, same files, every row of the manifest:
      * the manifest and the unit dictionary are parsed once;
      * rows are grouped by the frame count the manifest implies (<= max_batch per group; 16 utterances of 4 s already run at
        the GPU's full rate, and smaller groups pipeline finer: 19.5 k audio-s/s against 18.9 k with 32); io_threads host threads read
        the .npy files of the NEXT groups while the GPU works on the current one, and write the wav files of finished
        groups (the file I/O was 58 % of the job when done inline).  The per-file work itself is native: each of those
        threads hands a whole group to l2s_io_read_npy_f32 / l2s_io_write_wav_i16 (include/l2s_hand_off.h), which run it
        on native_threads host threads outside the interpreter (done per file in Python the job was interpreter-bound:
        16.5 k audio-s/s with 2 Python threads, 12.3 k with 16);
      * rows of a group whose exact lengths agree share a forward; inputs and int16 waveforms move through
        dispatch.HostPipeline (copies overlapped with the forward, int16 made on the device).
    `dataset_dir` replaces the absolute root in the manifest's first line (create_dataset.vocoder() writes the dataset
    where the service runs).  Returns the wav paths in manifest order; the bytes equal the per-utterance flow."""
    from concurrent.futures import ThreadPoolExecutor
    try:
        from .dispatch import HostPipeline
    except ImportError:                                   # top-level module use (see models_multi_input.py)
        from dispatch import HostPipeline
    manifest = os.path.join(dataset_dir, "label", split + ".tsv")
    _, rows = parse_manifest(manifest)
    code_dict = load_code_dict(code_dict_path or os.path.join(dataset_dir, "label", "dict.unt.txt"))
    dev = torch.device(device)
    paths: List[str] = []                                 # filled below, while the first groups are already loading
    num_mels = int(getattr(getattr(generator, "h", None), "num_mels", 80) or 80)

    def load_group(idxs):
        return _load_group(dataset_dir, rows, idxs, code_dict, pin=True, num_mels=num_mels, native_threads=native_threads)

    def write_group(done, wav, grp):
        done.synchronize()
        _write_group_native([paths[it[0]] for it in grp], wav, [it[4] for it in grp], native_threads)

    groups = _plan_groups(rows, max_batch, first_group)
    pipe = HostPipeline(generator, dev)
    with ThreadPoolExecutor(max_workers=max(2, io_threads)) as pool:
        loads = [pool.submit(load_group, g_) for g_ in groups]               # later groups load while earlier ones run
        paths.extend(os.path.join(out_dir, output_name(r) + ".wav") for r in rows)   # off the critical path: the loaders are native
        for d in {os.path.dirname(p) for p in paths}:
            os.makedirs(d, exist_ok=True)
        writers = []
        for fut in loads:
            for grp, code, mel, spk, wav in fut.result():
                done = pipe.submit(code, mel, spk, wav, mel_time_major=True)
                writers.append(pool.submit(write_group, done, wav, grp))     # wav files written while later groups run
        for w in writers:
            w.result()
    pipe.finish()
    generator.check_index_errors(dev)
    return paths
