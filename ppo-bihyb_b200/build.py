"""Build recipe of the CUDA library (in-tree, sm_100a only): ``python ppo-bihyb_b200/build.py``.

``csrc/ged_kernels.cu`` is compiled 9 times in parallel: once per slot count S (``-DGED_TU_S=S``: the solve / LSAPE / K*X
kernels of that size class), once for the CTA-per-candidate ("team") solve kernels (``-DGED_TU_TEAM``) and once for the
size-independent kernels, the dispatch and the C-ABI; the objects are linked into ``libged_b200.so``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ged_kernels.cu")
HDR = os.path.join(HERE, "..", "include", "ged_b200.h")
OUT = os.path.join(HERE, "libged_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SLOT_COUNTS = (1, 2, 3, 4, 5, 6, 7)

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false",                    # the canonical arithmetic has no FMA contraction (oracle/ged_oracle.c)
    "-Xcompiler", "-fPIC",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in (SRC, HDR, __file__))


def build(force: bool = False, verbose: bool = False, extra=(), out: str = None) -> str:
    """``out``: build a VARIANT library next to the product one (``libged_b200.<out>.so``, objects in ``build/<out>/``) for
    in-process A/B runs through ``GED_B200_LIBRARY``; the product library is only ever built without it."""
    if out is None and not (force or needs_build()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = OBJ_DIR if out is None else os.path.join(OBJ_DIR, out)
    target = OUT if out is None else os.path.join(HERE, f"libged_b200.{out}.so")
    os.makedirs(obj_dir, exist_ok=True)
    units = [("main", []), ("team", ["-DGED_TU_TEAM"])] + [(f"s{k}", [f"-DGED_TU_S={k}"]) for k in SLOT_COUNTS]

    def compile_unit(unit):
        name, defs = unit
        obj = os.path.join(obj_dir, f"ged_{name}.o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, SRC]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode:
            raise subprocess.CalledProcessError(res.returncode, cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_unit, units))
    subprocess.check_call([nvcc, "-shared", "-o", target] + objs)
    return target


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    variant = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    if extra and variant is None:
        raise SystemExit("-D flags build a variant: name it with --out=<name> (the product library is built without flags)")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra=extra, out=variant))
