"""ORACLE / TEST INFRASTRUCTURE.  Packs the reference's OWN driver scripts, byte for byte, into one archive under the
git-ignored oracle/_ref/ so that a GPU box (which has no /root/reference) can execute them against the drop-in:

    /root/reference/notebooks/{train_multimodal_fusion,extended_metrics}.py  ->  oracle/_ref/reference_drivers.tar.gz

oracle/_ref/ is listed in .gitignore (never enters history) and not in .gpurunignore (travels with the snapshot, like the
built .so).  Nothing is edited: MANIFEST.json records the sha256 of every member and `materialize()` checks it when
tests/test_gpu_reference_drivers.py unpacks the scripts into a temporary directory to run them.  Called by
__graft_entry__.build() when /root/reference exists; a no-op elsewhere.  The product never imports these files."""
import hashlib
import io
import json
import os
import shutil
import tarfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
STAGED = os.path.join(ROOT, "oracle", "_ref")
ARCHIVE = os.path.join(STAGED, "reference_drivers.tar.gz")
MANIFEST = os.path.join(STAGED, "MANIFEST.json")
DRIVERS = ["notebooks/train_multimodal_fusion.py", "notebooks/extended_metrics.py"]


def _sha(data):
    return hashlib.sha256(data).hexdigest()


def stage(reference=REFERENCE):
    """Returns {relative path: sha256} of the packed drivers, or None when the reference is not available here."""
    if not os.path.isdir(reference):
        return None
    os.makedirs(STAGED, exist_ok=True)
    for stale in ("notebooks",):                      # earlier layout: loose copies
        shutil.rmtree(os.path.join(STAGED, stale), ignore_errors=True)
    manifest = {}
    with tarfile.open(ARCHIVE, "w:gz") as tar:
        for rel in DRIVERS:
            with open(os.path.join(reference, rel), "rb") as f:
                data = f.read()
            manifest[rel] = _sha(data)
            info = tarfile.TarInfo(rel)
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))
    with open(MANIFEST, "w") as f:
        json.dump({"source": reference, "sha256": manifest}, f, indent=1)
    return manifest


def materialize(dest):
    """Writes the driver scripts below `dest` (dest/notebooks/*.py), from /root/reference when present, else from the staged
    archive (sha256 of every member checked against MANIFEST.json).  Returns `dest`, or None when neither source exists."""
    if all(os.path.exists(os.path.join(REFERENCE, rel)) for rel in DRIVERS):
        for rel in DRIVERS:
            os.makedirs(os.path.dirname(os.path.join(dest, rel)), exist_ok=True)
            shutil.copyfile(os.path.join(REFERENCE, rel), os.path.join(dest, rel))
        return dest
    if not (os.path.exists(ARCHIVE) and os.path.exists(MANIFEST)):
        return None
    with open(MANIFEST) as f:
        want = json.load(f)["sha256"]
    with tarfile.open(ARCHIVE, "r:gz") as tar:
        for rel in DRIVERS:
            data = tar.extractfile(rel).read()
            if _sha(data) != want[rel]:
                raise RuntimeError("staged reference driver does not match its recorded sha256: " + rel)
            os.makedirs(os.path.dirname(os.path.join(dest, rel)), exist_ok=True)
            with open(os.path.join(dest, rel), "wb") as out:
                out.write(data)
    return dest


if __name__ == "__main__":
    print(stage())
