"""Builds libtpb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

csrc/tpb200.cu is compiled seven times in parallel: once per (ndims, eltype, coordinates eltype)
combination (`-DTPB_TU_TAG=...`: the kernels of that combination behind five entry functions)
and once as the main unit (dispatch + C ABI); the objects are linked into one shared library.
One unit with all ~150 kernel instantiations takes 3.5 minutes; `-split-compile` halves that
but costs 7.7 % in k_interact_tiles (0.942 vs 0.875 ms, measured on B200), so it is not used."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libtpb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xfatbin", "-compress-all",
]

# (tag, ndims, eltype, coordinates eltype); None = the main unit
UNITS = [None, ("3ff", 3, "float", "float"), ("3fd", 3, "float", "double"), ("3dd", 3, "double", "double"),
         ("2ff", 2, "float", "float"), ("2fd", 2, "float", "double"), ("2dd", 2, "double", "double")]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "tpb200.h")]


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def _source_stamp() -> str:
    import hashlib
    h = hashlib.sha1()
    for src in sources():
        with open(src, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False, out: str = SO, unit_flags=None) -> str:
    """Compile the translation units (objects cached under build/obj by source hash + flags, so a
    tuning variant that only adds -D flags to one unit recompiles that unit alone) and link them.
    `unit_flags`: {unit tag: [extra nvcc flags]} for variant builds (tools/build_variant.py)."""
    if not force and out == SO and not unit_flags and not is_stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("TPB_NVCC_EXTRA", "").split()
    src = os.path.join(CSRC, "tpb200.cu")
    cache = os.path.join(os.path.dirname(HERE), "build", "obj")
    os.makedirs(cache, exist_ok=True)
    stamp = _source_stamp()
    unit_flags = unit_flags or {}

    def compile_unit(unit):
        import hashlib
        name = "main" if unit is None else unit[0]
        cmd = [nvcc, *NVCC_FLAGS, *extra, *unit_flags.get(name, [])]
        if unit is not None:
            tag, nd, t, ct = unit
            cmd += [f"-DTPB_TU_TAG={tag}", f"-DTPB_TU_ND={nd}", f"-DTPB_TU_T={t}", f"-DTPB_TU_CT={ct}"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        key = hashlib.sha1((stamp + " ".join(cmd)).encode()).hexdigest()[:16]
        obj = os.path.join(cache, f"tpb200_{name}_{key}.o")
        if os.path.exists(obj) and not verbose and not (force and not unit_flags):
            return obj, ""
        tmp_obj = obj + f".{os.getpid()}.tmp"
        res = subprocess.run(cmd + ["-c", "-o", tmp_obj, src], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed ({name}):\n" + res.stdout + res.stderr)
        os.replace(tmp_obj, obj)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    if verbose:
        for _, log in results:
            print(log)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [o for o, _ in results]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    # keep the cache small: drop objects of other source states
    for f in os.listdir(cache):
        path = os.path.join(cache, f)
        if f.endswith(".o") and path not in {o for o, _ in results} and os.path.getmtime(path) < os.path.getmtime(out) - 6 * 3600:
            os.unlink(path)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
