"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference modules of the hot path for the CPU baseline.

`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own `src.egnn.Dynamics.forward`
(/root/reference/src/egnn.py:374-447) on the GPU box's host cores. /root/reference does not exist there, so this
recipe (run by `__graft_entry__.build()` in the build container, where it does) copies the four pure torch/numpy modules
that call needs -- egnn.py, utils.py, noise.py, edm.py; none of them imports rdkit / pytorch_lightning -- byte for byte into
`oracle/_ref/src/`, together with a manifest of their sha256 sums. `oracle/_ref/` is git-ignored (nothing of the reference
enters the history) but is not in `.gpurunignore`, so it travels to the GPU box like the built `.so`.
When `oracle/_ref/` is absent, bench.py falls back to the oracle port (`kind: "port"`).
"""
import hashlib
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DIFFLINKER_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
MODULES = ["egnn.py", "utils.py", "noise.py", "edm.py"]


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DEST, "src", m)) for m in MODULES)


def build_ref(verbose: bool = True) -> bool:
    src_dir = os.path.join(REF_ROOT, "src")
    if not all(os.path.isfile(os.path.join(src_dir, m)) for m in MODULES):
        if verbose:
            print(f"oracle/_ref: reference not present at {REF_ROOT}; keeping what is staged ({'yes' if staged() else 'nothing'})")
        return staged()
    os.makedirs(os.path.join(DEST, "src"), exist_ok=True)
    manifest = {"source": REF_ROOT, "files": {}}
    try:
        manifest["commit"] = subprocess.run(["git", "-C", REF_ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except OSError:
        manifest["commit"] = None
    for m in MODULES:
        shutil.copyfile(os.path.join(src_dir, m), os.path.join(DEST, "src", m))
        with open(os.path.join(src_dir, m), "rb") as f:
            manifest["files"][m] = hashlib.sha256(f.read()).hexdigest()
    open(os.path.join(DEST, "src", "__init__.py"), "w").close()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f"oracle/_ref: staged {', '.join(MODULES)} from {REF_ROOT}")
    return True


if __name__ == "__main__":
    build_ref()
