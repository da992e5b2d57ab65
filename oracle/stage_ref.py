"""Stage the UNMODIFIED reference sources this path needs under oracle/_ref/ (test infrastructure).

The reference is a pure-Python script tree (no build system): its quantizer modules are `models/vqvae.py:10-259`
and its training driver is `scripts/train_ablation.py`.  `/root/reference` exists only in the build container, so
`__graft_entry__.build()` runs this recipe there; the copies travel to the GPU box with the snapshot (oracle/_ref/ is
git-ignored, NOT gpurun-ignored -- exactly like the built libvqb200.so) and are used as

  * the CPU baseline / `bench.py --impl reference` arm (`kind: "reference"`): the reference's own torch modules on the
    box's host cores;
  * the driver of `tests/test_gpu_dropin.py::test_unmodified_train_ablation_script`: the reference's own training
    script, byte-identical, run against this repo's drop-in `models/vqvae.py`.

Nothing under oracle/_ref/ is ever imported by the product path.  Files are copied verbatim; MANIFEST.json records
their sha256 so a test can show they were not edited.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/vqvae.py", "models/experiment_config.py", "scripts/train_ablation.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(reference_root="/root/reference", dest=DEST):
    """Copy FILES from the reference tree; returns the manifest (or None when the reference is absent)."""
    if not os.path.isdir(reference_root):
        return None
    manifest = {"source": reference_root, "files": {}}
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = sha256(dst)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def staged(rel):
    """Path of a staged file, or None."""
    p = os.path.join(DEST, rel)
    return p if os.path.exists(p) else None


def load_reference_vqvae():
    """Import the staged, unmodified `models/vqvae.py` under a private module name (the repo root has its own
    `models` package: the drop-in).  Raises FileNotFoundError when oracle/_ref has not been staged."""
    import importlib.util
    path = staged("models/vqvae.py")
    if path is None:
        raise FileNotFoundError("oracle/_ref/models/vqvae.py is not staged (run `python oracle/stage_ref.py` where "
                                "/root/reference exists)")
    spec = importlib.util.spec_from_file_location("_reference_vqvae", path)
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = stage(*(sys.argv[1:2]))
    print(json.dumps(m, indent=1) if m else "reference tree not found: nothing staged")
