"""ORACLE / TEST INFRASTRUCTURE.  Stages the reference's OWN driver scripts, byte for byte, into the git-ignored
oracle/_ref/ so that a GPU box (which has no /root/reference) can execute them against the drop-in:

    /root/reference/notebooks/train_multimodal_fusion.py  ->  oracle/_ref/notebooks/train_multimodal_fusion.py
    /root/reference/notebooks/extended_metrics.py         ->  oracle/_ref/notebooks/extended_metrics.py

oracle/_ref/ is listed in .gitignore (never enters history) and not in .gpurunignore (travels with the snapshot, like the
built .so).  Nothing is edited: tests/test_gpu_reference_drivers.py checks the sha256 recorded here before running them.
Called by __graft_entry__.build() when /root/reference exists; a no-op elsewhere.  The product never imports these files."""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
STAGED = os.path.join(ROOT, "oracle", "_ref")
DRIVERS = ["notebooks/train_multimodal_fusion.py", "notebooks/extended_metrics.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(reference=REFERENCE, dest=STAGED):
    """Returns {relative path: sha256} of the staged drivers, or None when the reference is not available here."""
    if not os.path.isdir(reference):
        return None
    manifest = {}
    for rel in DRIVERS:
        src, dst = os.path.join(reference, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
        assert manifest[rel] == _sha(src)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "sha256": manifest}, f, indent=1)
    return manifest


def locate():
    """Directory holding notebooks/*.py of the reference: /root/reference if present, else the staged copy, else None."""
    for base in (REFERENCE, STAGED):
        if all(os.path.exists(os.path.join(base, rel)) for rel in DRIVERS):
            return base
    return None


if __name__ == "__main__":
    print(stage())
