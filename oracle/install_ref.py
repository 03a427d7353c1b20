"""TEST / BENCH INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference importable on the GPU box.

The reference is a plain-Python tree without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` does not apply; the equivalent here is a verbatim
copy of the `switch_nerf` package into the git-ignored `baseline/_ref/` (it travels to the GPU box with
the gpurun snapshot, it is never committed).  `/root/reference` exists only in the build container:
`__graft_entry__.build()` calls `install()` there; on the GPU box the copy is simply used.

Users: oracle/make_golden_cuda.py (CUDA-autocast goldens of the unmodified reference), bench.py's
`gpu_comparator` leg, tests/test_reference_in_loop.py.  Nothing under switch_nerf_b200/ imports it.
"""
import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/switch_nerf"
DST_ROOT = os.path.join(ROOT, "baseline", "_ref")
DST = os.path.join(DST_ROOT, "switch_nerf")


def reference_root():
    """Directory to put on sys.path so that `import switch_nerf` finds the unmodified reference (or None)."""
    if os.path.isdir(SRC):
        return os.path.dirname(SRC)
    if os.path.isdir(DST):
        return DST_ROOT
    return None


def install(verbose=False):
    """Copy /root/reference/switch_nerf -> baseline/_ref/switch_nerf (python sources + yaml configs only)."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    n = 0
    for dirpath, dirnames, filenames in os.walk(SRC):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC)
        out = os.path.join(DST, rel) if rel != "." else DST
        os.makedirs(out, exist_ok=True)
        for f in filenames:
            if not f.endswith((".py", ".yaml", ".yml")):
                continue
            s, d = os.path.join(dirpath, f), os.path.join(out, f)
            if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
                shutil.copyfile(s, d)
                n += 1
    if verbose:
        print(f"baseline/_ref: {n} files refreshed from {SRC}")
    return True


if __name__ == "__main__":
    install(verbose=True)
