"""Build libiins_b200.so (sm_100a) in-tree with nvcc.  No torch headers: the library is a plain
C-ABI shared object (include/iins_b200.h) that Python binds with ctypes.

Every csrc/*.cu is one translation unit, compiled in parallel into csrc/_obj/*.o (rebuilt when the source or any header
is newer) and linked into the shared object.  The link goes to a temporary file that is renamed into place under a file
lock, so concurrent ranks of a torchrun launch never see a half-written library."""
import fcntl
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libiins_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".inc"))] + [
        os.path.join(os.path.dirname(HERE), "include", "iins_b200.h")]


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(OUT, _sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():             # another process built it while we waited
            return OUT
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        hdrs = _headers()

        def compile_one(src):
            obj = _obj(src)
            if not force and not _stale(obj, [src] + hdrs):
                return ""
            cmd = [nvcc] + FLAGS + ["-Xptxas", "-v" if verbose else "-warn-spills", "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{r.stdout}{r.stderr}")
            return r.stderr

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            logs = list(pool.map(compile_one, _sources()))
        if verbose:
            sys.stderr.write("".join(logs))
        tmp = OUT + f".tmp{os.getpid()}"
        r = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [_obj(s) for s in _sources()] +
                           ["-o", tmp], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libiins_b200.so")
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
