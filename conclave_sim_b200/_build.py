"""Build recipe of the CUDA library (nvcc, sm_100a only).  Used by __graft_entry__.build().

One object per translation unit of csrc/ (cabi.cu + one eng_*.cu per engine), compiled in parallel and linked into
conclave_sim_b200/libconclave_b200.so; an object is rebuilt when its source or any header is newer than it."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libconclave_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(ROOT, "include", "conclave_b200.h")]


def sources():
    return units() + headers()


def _obj(unit: str) -> str:
    return os.path.join(OBJ, os.path.basename(unit)[:-3] + ".o")


def _stale(unit: str) -> bool:
    o = _obj(unit)
    return not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s) for s in [unit] + headers())


def up_to_date() -> bool:
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in sources())


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    todo = [u for u in units() if force or _stale(u)]

    def compile_one(unit):
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", unit, "-o", _obj(unit)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        logs = list(pool.map(compile_one, todo))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC"] + [_obj(u) for u in units()] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB
