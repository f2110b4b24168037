"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe that makes the UNMODIFIED reference importable on the GPU box.

/root/reference exists only in the build container.  `python -m oracle.build_ref` copies the reference's own Python package
(`opencood/**/*.py` and its yaml hypes, byte for byte, nothing edited) into `oracle/_ref/`, which is git-ignored (reference
sources never enter this repository's history) but not gpurun-ignored, so it travels to the GPU box next to the built .so
files.  `oracle/ref_shim.py` then imports it from there.  `__graft_entry__.build()` runs this when /root/reference is present.

Users: `bench.py --impl reference` (the reference's own CPU path), bench.py's `cuda_eager_reference` baseline leg (the same
modules on the GPU through PyTorch/cuDNN) and the boundary tests that drive `opencood.tools.train_utils.create_model`.
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get("HEAL_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yaml", ".pyx")
SKIP_DIRS = {"__pycache__", "logs"}


def build(verbose: bool = False) -> str:
    src_pkg = os.path.join(SRC, "opencood")
    if not os.path.isdir(src_pkg):
        if os.path.isdir(os.path.join(DST, "opencood")):
            return DST                      # GPU box: use what travelled with the snapshot
        raise RuntimeError(f"reference not found at {SRC} and no prebuilt {DST}")
    n, h = 0, hashlib.sha256()
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        rel = os.path.relpath(root, SRC)
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            out_dir = os.path.join(DST, rel)
            os.makedirs(out_dir, exist_ok=True)
            shutil.copyfile(os.path.join(root, f), os.path.join(out_dir, f))
            with open(os.path.join(root, f), "rb") as fh:
                h.update(fh.read())
            n += 1
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as fh:
        fh.write(f"unmodified copy of {SRC}/opencood ({n} files, sha256 of the concatenation {h.hexdigest()})\n")
    if verbose:
        print(f"[oracle.build_ref] {n} files -> {DST}")
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
