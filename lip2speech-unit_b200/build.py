"""Builds the C-ABI CUDA library in-tree (lib/libl2s_vocoder.so) for sm_100a.

    python lip2speech-unit_b200/build.py [--force] [-v]

One object per translation unit under csrc/ (the tcgen05 kernel families each live in their own .cu so that they
compile in parallel and an edit rebuilds only what depends on it: dependencies come from nvcc -MD), then one
link.  nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot (it is
git-ignored, not gpurun-ignored).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libl2s_vocoder.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _deps_of(src):
    """Files the object depends on, from the nvcc -MD output of its last build (None: never built)."""
    dep = _obj(src)[:-2] + ".d"
    if not os.path.exists(dep) or not os.path.exists(_obj(src)):
        return None
    with open(dep) as f:
        text = f.read().replace("\\\n", " ")
    files = text.split(":", 1)[1].split() if ":" in text else []
    return [p for p in files if not p.startswith(("/usr/", "/opt/"))] + [src, os.path.abspath(__file__)]


def _digest(src, deps):
    """Content hash of everything the object depends on (sources, headers, flags, this script): a copied tree whose
    modification times were not preserved (a snapshot sent to the GPU box, a fresh checkout next to cached objects) must
    not trigger -- or miss -- a rebuild."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    # the .d file holds the absolute paths of the build that wrote it; the tree may since have moved (the GPU box runs a
    # copy under another root): files are looked up by NAME under csrc/ and include/
    names = sorted(set(os.path.basename(x) for x in deps))
    for name in names:
        h.update(name.encode())
        for base in (CSRC, os.path.join(HERE, "..", "include"), HERE):
            path = os.path.join(base, name)
            if os.path.exists(path):
                with open(path, "rb") as f:
                    h.update(f.read())
                break
        else:
            h.update(b"<missing>")
    return h.hexdigest()


def _stamp(src):
    return _obj(src)[:-2] + ".stamp"


def _stale(src) -> bool:
    deps = _deps_of(src)
    if deps is None or not os.path.exists(_stamp(src)):
        return True
    with open(_stamp(src)) as f:
        return f.read().strip() != _digest(src, deps)


def _lib_stamp():
    return os.path.join(OBJ_DIR, "lib.stamp")


def _lib_digest():
    h = hashlib.sha256()
    for s_ in _sources():
        with open(_stamp(s_)) as f:
            h.update(f.read().encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(_lib_stamp()):
        return True
    if any(_stale(s) for s in _sources()):
        return True
    with open(_lib_stamp()) as f:
        return f.read().strip() != _lib_digest()


def _compile(src, nvcc, verbose):
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-MD", "-MF", _obj(src)[:-2] + ".d", "-c", "-o", _obj(src), src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, cmd, proc


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in _sources() if force or _stale(s)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        for src, cmd, proc in pool.map(lambda s: _compile(s, nvcc, verbose), todo):
            if proc.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
            with open(_stamp(src), "w") as f:
                f.write(_digest(src, _deps_of(src)))
            if verbose:
                sys.stderr.write(proc.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [_obj(s) for s in _sources()]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    with open(_lib_stamp(), "w") as f:
        f.write(_lib_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
