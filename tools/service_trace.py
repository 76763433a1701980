"""Where the time of one serve_vocoder_request goes (host time stamps around the same calls the service makes).
python tools/service_trace.py [n_utts]"""
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import vocoder_oracle as vo  # noqa: E402

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pkg = ge.load_package()
ho = pkg.hand_off
dev = torch.device("cuda:0")
h = vo.shipped_config()
g = pkg.MelCodeGenerator(pkg.AttrDict(h))
g.load_state_dict(vo.init_state_dict(h, seed=1234, style="ref"), strict=True)
g.eval(); g.remove_weight_norm(); g = g.to(dev)
print("host cores:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))

with tempfile.TemporaryDirectory() as root:
    os.makedirs(os.path.join(root, "label"))
    with open(os.path.join(root, "label", "dict.unt.txt"), "w") as f:
        f.writelines(f"{i} 1\n" for i in range(200))
    lengths = [400, 400, 400, 300, 400, 200, 400, 300]
    tsv, unt, total_s = [root + "\n"], [], 0.0
    rng = np.random.default_rng(0)
    for i in range(n_utts):
        frames = lengths[i % len(lengths)]
        rel = f"audio/test/spk{i % 4}/{i:05d}.wav"
        for sub, arr in (("mel", rng.standard_normal((frames, 80)).astype(np.float32)), ("spk_emb", rng.standard_normal(256).astype(np.float32))):
            path = os.path.join(root, rel.replace("audio/", sub + "/")[:-4] + ".npy")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.save(path, arr)
        tsv.append(f"test/spk{i % 4}/{i:05d}\tvideo/x.mp4\t{rel}\t{frames // 4}\t{frames * 160}\n")
        unt.append(" ".join(str(int(c)) for c in rng.integers(0, 200, frames // 2)) + "\n")
        total_s += frames / 100.0
    open(os.path.join(root, "label", "test.tsv"), "w").writelines(tsv)
    open(os.path.join(root, "label", "test.unt"), "w").writelines(unt)

    def traced(out_dir, io_threads=2, native_threads=4, max_batch=32, first_group=8):
        T = [("start", time.perf_counter())]
        mark = lambda name: T.append((name, time.perf_counter()))      # noqa: E731
        _, rows = ho.parse_manifest(os.path.join(root, "label", "test.tsv"))
        code_dict = ho.load_code_dict(os.path.join(root, "label", "dict.unt.txt"))
        mark("parsed")
        paths = [os.path.join(out_dir, ho.output_name(r) + ".wav") for r in rows]
        for d in {os.path.dirname(p) for p in paths}:
            os.makedirs(d, exist_ok=True)
        groups = ho._plan_groups(rows, max_batch, first_group)
        mark("planned+mkdirs")
        pipe = pkg.dispatch.HostPipeline(g, dev)
        mark("pipeline")
        load_t, write_t = [], []

        def load_group(idxs):
            t0 = time.perf_counter()
            r = ho._load_group(root, rows, idxs, code_dict, pin=True, native_threads=native_threads)
            load_t.append((len(idxs), t0 - T[0][1], time.perf_counter() - T[0][1]))
            return r

        def write_group(done, wav, grp):
            t0 = time.perf_counter()
            done.synchronize()
            t1 = time.perf_counter()
            ho._write_group_native([paths[it[0]] for it in grp], wav, [it[4] for it in grp], native_threads)
            write_t.append((len(grp), t0 - T[0][1], t1 - T[0][1], time.perf_counter() - T[0][1]))

        with ThreadPoolExecutor(max_workers=max(2, io_threads)) as pool:
            loads = [pool.submit(load_group, g_) for g_ in groups]
            writers = []
            for k, fut in enumerate(loads):
                res = fut.result()
                mark(f"group{k} loaded")
                for grp, code, mel, spk, wav in res:
                    done = pipe.submit(code, mel, spk, wav, mel_time_major=True)
                    writers.append(pool.submit(write_group, done, wav, grp))
                mark(f"group{k} submitted")
            for w in writers:
                w.result()
            mark("writers done")
        pipe.finish()
        g.check_index_errors(dev)
        mark("finished")
        return T, load_t, write_t

    for rep in range(4):
        with tempfile.TemporaryDirectory() as out:
            torch.cuda.synchronize()
            T, load_t, write_t = traced(out)
            torch.cuda.synchronize()
    t0 = T[0][1]
    print(f"total {1e3 * (T[-1][1] - t0):.2f} ms -> {total_s / (T[-1][1] - t0):.0f} audio-s/s")
    for name, t in T:
        print(f"  {1e3 * (t - t0):8.2f} ms  {name}")
    print("loads (rows, start, end ms):", [(n, round(1e3 * a, 2), round(1e3 * b, 2)) for n, a, b in load_t])
    print("writes (rows, start, gpu done, end ms):", [(n, round(1e3 * a, 2), round(1e3 * b, 2), round(1e3 * c, 2)) for n, a, b, c in write_t])
