"""TEST INFRASTRUCTURE ONLY -- recipe for ``oracle/_ref/``: the reference's own lower-level GED solver, so that the CPU arm of
``bench.py`` (``--impl reference``, ``cpu_baseline``) can time the REFERENCE instead of the C port on the GPU box.

    python oracle/build_ref.py          (build container only: needs /root/reference)

The reference's path is pure Python (``utils/ged_env.py`` + ``utils/sinkhorn.py``, imported by it): there is nothing to
compile, so "building" is placing the two files, byte for byte, under ``oracle/_ref/utils/`` with a manifest of their SHA-256
sums.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never enter the history) and NOT in ``.gpurunignore``
(it travels to the GPU box like a built ``.so``).  ``oracle/ref_shim.py`` imports it from there when ``/root/reference`` is
absent; the third-party imports it needs beyond torch / SciPy / numpy (``torch_geometric``, the Cython ``a_star``) are the
shim's stand-ins.  Nothing under ``ppo-bihyb_b200/`` may import any of this.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("PPO_BIHYB_REFERENCE", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("utils/ged_env.py", "utils/sinkhorn.py")


def available() -> bool:
    return all(os.path.isfile(os.path.join(DST_ROOT, f)) for f in FILES)


def build(verbose: bool = True) -> bool:
    """Returns True when ``oracle/_ref`` is in place afterwards (already there, or copied now)."""
    if not all(os.path.isfile(os.path.join(SRC_ROOT, f)) for f in FILES):
        return available()                  # the GPU box: use what travelled with the snapshot
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(SRC_ROOT, f), os.path.join(DST_ROOT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[f] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest,
                   "note": "verbatim copies of the reference's files; git-ignored; test infrastructure only"}, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files placed (sha256 in MANIFEST.json)")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
