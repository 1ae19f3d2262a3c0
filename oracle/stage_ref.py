"""Stage the UNMODIFIED reference under oracle/_ref/ so that it can be imported where /root/reference does not
exist (the GPU box).  Test / measurement infrastructure only -- nothing under drakegpt_b200/ may import it.

    python oracle/stage_ref.py            # run in the build container (needs /root/reference)

Copies the four pure-Python source files of the hot path byte for byte (src/model.py, src/model_component.py,
src/config.py, src/preprocessing.py) into oracle/_ref/src/ and records their SHA-256 next to them.  oracle/_ref/ is
git-ignored (no reference source ever enters the history) but travels to the GPU box with the gpurun snapshot, like
the built .so files.  `bench.py --impl reference` and the `cpu_baseline` leg import it from there through
``load_reference()``; when it is absent they fall back to the oracle port and say so (`kind: "port"`).
"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "src")
FILES = ("model.py", "model_component.py", "config.py", "preprocessing.py")


def stage(ref_root=None):
    ref_root = ref_root or os.environ.get("DRAKE_REF", "/root/reference")
    src = os.path.join(ref_root, "src")
    if not os.path.isdir(src):
        return False
    os.makedirs(DEST, exist_ok=True)
    sums = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DEST, f))
        with open(os.path.join(DEST, f), "rb") as fh:
            sums[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "SHA256.json"), "w") as fh:
        json.dump(sums, fh, indent=1)
    return True


def available():
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def load_reference():
    """Import the staged reference's model module (its model_component import resolves inside oracle/_ref/src)."""
    if not available():
        return None
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    mods = {}
    for name in ("model_component", "model"):
        spec = importlib.util.spec_from_file_location("_drake_ref_" + name, os.path.join(DEST, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        if name == "model_component":
            sys.modules.setdefault("model_component", mod)  # `from model_component import ...` inside model.py
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["model"]


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "reference not found", DEST)
