"""Recipe that makes the UNMODIFIED reference travel to the GPU box: ``oracle/_ref/reference``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python, so "building" it means
taking a byte-for-byte snapshot of the files the hot path and its callers live in.  The snapshot is a
BUILD OUTPUT, exactly like ``librvb.so``: ``oracle/_ref/`` is git-ignored (no reference source ever enters
the history) but not gpurun-ignored, so the copy made by ``__graft_entry__.build()`` in the build
container is present on the GPU box, where ``/root/reference`` does not exist.

What uses it (and nothing else may): ``oracle/reference_loader.py`` -> the ``-m gpu`` tests that run the
reference's own ``UNet`` / ``UNet_Onset`` / ``OnsetsAndFrames_VAT_full`` patched and unpatched on the same
B200 (tests/test_gpu_reference_models.py), ``bench.py --impl reference`` (``kind: "reference"``) and
``bench.py``'s ``gpu_eager_baseline`` leg (the reference's eager PyTorch path on the same GPU).

    python -m oracle.ref_snapshot          # (re)create oracle/_ref/reference + MANIFEST.json
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("RECONVAT_REFERENCE_SRC", "/root/reference")
DEST = os.path.join(HERE, "_ref", "reference")

# the model package (hot path + its callers) and the scripts that construct the models
FILES = [
    "model/__init__.py", "model/constants.py", "model/utils.py", "model/Spectrogram.py", "model/VAT.py",
    "model/self_attention_VAT.py", "model/UNet_onset.py", "model/onset_frame_VAT.py", "model/Segmentation.py",
    "model/decoding.py", "model/helper_functions.py", "model/dataset.py", "model/evaluate_functions.py",
    "model/midi.py", "train_UNet_VAT.py", "train_UNet_Onset_VAT.py", "train_baseline_onset_frame_VAT.py",
    "transcribe_files.py", "requirements.txt",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def materialise(source=SOURCE, dest=DEST, quiet=False):
    """Copy FILES from ``source`` to ``dest`` (byte for byte) and write MANIFEST.json with their sha256.
    Returns the manifest, or None when the reference tree is not there (GPU box: the shipped copy is used)."""
    if not os.path.isdir(os.path.join(source, "model")):
        return None
    manifest = {"source": source, "files": {}}
    for rel in FILES:
        src = os.path.join(source, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = _sha(dst)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print("oracle/_ref: %d reference files snapshotted from %s" % (len(manifest["files"]), source))
    return manifest


def verify(dest=DEST):
    """True when every file of the shipped snapshot still has the sha256 recorded at build time."""
    try:
        with open(os.path.join(dest, "MANIFEST.json")) as f:
            manifest = json.load(f)
    except OSError:
        return False
    return all(os.path.exists(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == digest
               for rel, digest in manifest["files"].items())


if __name__ == "__main__":
    sys.exit(0 if materialise() is not None else 1)
