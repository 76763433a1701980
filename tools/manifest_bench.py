"""Service-style throughput of the manifest caller: a synthetic dataset in the reference's on-disk
layout (label/test.tsv + .unt + dict, mel/*.npy, spk_emb/*.npy) is vocoded to pred_wav/*.wav

  (a) the way inference.py does it: one utterance per forward, host-side * 32768 -> int16, one wav per item;
  (b) hand_off.vocode_manifest: equal-length rows stacked into one forward, int16 on the device.

Both include reading the .npy files and writing the wav files.  Prints JSON.   python tools/manifest_bench.py [n_utts]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402  (weights + synthetic inputs)

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pkg = ge.load_package()
ho = pkg.hand_off
dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)

with tempfile.TemporaryDirectory() as root:
    os.makedirs(os.path.join(root, "label"))
    with open(os.path.join(root, "label", "dict.unt.txt"), "w") as f:
        f.writelines(f"{i} 1\n" for i in range(200))
    lengths = [400, 400, 400, 300, 400, 200, 400, 300]          # mel frames: mostly 4 s, some shorter
    tsv, unt, total_s = [root + "\n"], [], 0.0
    for i in range(n_utts):
        frames = lengths[i % len(lengths)]
        code, mel, spkr = vo.synthetic_inputs(1, frames, seed=1000 + i)
        rel = f"audio/test/spk{i % 4}/{i:05d}.wav"
        for sub, arr in (("mel", mel[0].T.numpy().astype(np.float32)), ("spk_emb", spkr[0].numpy().astype(np.float32))):
            path = os.path.join(root, rel.replace("audio/", sub + "/")[:-4] + ".npy")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.save(path, arr)
        tsv.append(f"test/spk{i % 4}/{i:05d}\tvideo/x.mp4\t{rel}\t{frames // 4}\t{frames * 160}\n")
        unt.append(" ".join(str(int(c)) for c in code[0]) + "\n")
        total_s += frames / 100.0
    open(os.path.join(root, "label", "test.tsv"), "w").writelines(tsv)
    open(os.path.join(root, "label", "test.unt"), "w").writelines(unt)
    manifest = os.path.join(root, "label", "test.tsv")

    def per_utterance(out_dir):
        _, rows = ho.parse_manifest(manifest)
        cd = ho.load_code_dict(os.path.join(root, "label", "dict.unt.txt"))
        for r in rows:
            feats, n = ho.load_item(root, r, cd)
            y = g(**{k: torch.from_numpy(v).to(dev).unsqueeze(0) for k, v in feats.items()})
            audio = (y.squeeze() * 32768.0).cpu().numpy().astype("int16")      # inference.py:79-81
            ho.write_wav_int16(os.path.join(out_dir, ho.output_name(r) + ".wav"), audio[:n])

    def batched(out_dir):
        ho.vocode_manifest(g, manifest, out_dir, root=root, device=dev)

    def service(out_dir, io_threads=None):
        ho.serve_vocoder_request(g, root, out_dir, device=dev, **({} if io_threads is None else {"io_threads": io_threads}))

    res = {"utterances": n_utts, "audio_s": total_s}
    for name, fn in (("per_utterance_like_inference_py", per_utterance), ("batched_vocode_manifest", batched),
                     ("serve_vocoder_request", service)):
        with tempfile.TemporaryDirectory() as out:
            fn(out)                                   # warm-up: plans, graphs, file cache
        with tempfile.TemporaryDirectory() as out:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            fn(out)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[name] = {"seconds": round(dt, 4), "audio_s_per_s": round(total_s / dt, 1)}
    sweep = {}
    for nt in (2, 3, 4, 6, 8, 16):
        with tempfile.TemporaryDirectory() as out:
            service(out, nt)
        best = 1e9
        for _ in range(3):
            with tempfile.TemporaryDirectory() as out:
                torch.cuda.synchronize(); t0 = time.perf_counter()
                service(out, nt)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        sweep[str(nt)] = round(total_s / best, 1)
    res["serve_vocoder_request_io_threads_sweep_audio_s_per_s"] = sweep
    sweep2 = {}
    for mb, fg, nt in ((32, 8, 2), (64, 8, 2), (64, 16, 2), (32, 4, 2), (16, 8, 2), (64, 8, 3), (32, 32, 2)):
        def run(out):
            ho.serve_vocoder_request(g, root, out, device=dev, max_batch=mb, first_group=fg, io_threads=nt)
        with tempfile.TemporaryDirectory() as out:
            run(out)
        best = 1e9
        for _ in range(4):
            with tempfile.TemporaryDirectory() as out:
                torch.cuda.synchronize(); t0 = time.perf_counter()
                run(out)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        sweep2[f"max_batch={mb},first={fg},threads={nt}"] = round(total_s / best, 1)
    res["serve_vocoder_request_sweep_audio_s_per_s"] = sweep2
    sweep3 = {}
    for nt, nn in ((2, 1), (2, 2), (2, 4), (2, 8), (3, 4), (4, 4), (3, 8)):
        def run(out):
            ho.serve_vocoder_request(g, root, out, device=dev, io_threads=nt, native_threads=nn)
        with tempfile.TemporaryDirectory() as out:
            run(out)
        best = 1e9
        for _ in range(4):
            with tempfile.TemporaryDirectory() as out:
                torch.cuda.synchronize(); t0 = time.perf_counter()
                run(out)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        sweep3[f"io_threads={nt},native_threads={nn}"] = round(total_s / best, 1)
    res["serve_vocoder_request_native_threads_sweep_audio_s_per_s"] = sweep3
    # host-only cost of the file I/O of the whole request, native batch calls against one numpy / Python call per file
    _, rows = ho.parse_manifest(manifest)
    cd = ho.load_code_dict(os.path.join(root, "label", "dict.unt.txt"))
    groups = ho._plan_groups(rows, 32, 8)
    host = {}
    for name, fn in (("native", ho._load_group), ("python", ho._load_group_python)):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            loaded = [fn(root, rows, g_, cd, pin=True) for g_ in groups]
            best = min(best, time.perf_counter() - t0)
        host["load_ms_" + name] = round(best * 1e3, 2)
    with tempfile.TemporaryDirectory() as out:
        paths = [os.path.join(out, f"{i}.wav") for i in range(len(rows))]
        t0 = time.perf_counter()
        for lg in loaded:
            for grp, code, mel, spk, wav in lg:
                ho._write_group_native([paths[it[0]] for it in grp], wav, [it[4] for it in grp], 4)
        host["write_ms_native"] = round((time.perf_counter() - t0) * 1e3, 2)
        t0 = time.perf_counter()
        for lg in loaded:
            for grp, code, mel, spk, wav in lg:
                for k, it in enumerate(grp):
                    ho.write_wav_int16(paths[it[0]], wav[k, :it[4]].numpy())
        host["write_ms_python"] = round((time.perf_counter() - t0) * 1e3, 2)
    res["host_file_io_of_the_request"] = host
    print(json.dumps(res))
