"""ORACLE-side recipe (test / baseline infrastructure, never imported by the product path): copy the reference's own
source files for the hot path, VERBATIM, from /root/reference into the git-ignored ``baseline/_ref/`` so that the
reference arm of bench.py (``--impl reference``) and the informational GPU "library bar" can run the reference's code
on the GPU box, where /root/reference does not exist (gpurun ships /root/repo only; ``baseline/_ref/`` is git-ignored
but not gpurun-ignored).  Nothing is edited; ``MANIFEST.json`` records the sha256 of every file.  The reference is
pure Python with no setup.py / pyproject.toml, so this copy is the whole "install" (DESIGN.md §7).

    python -m oracle.make_ref            # run by __graft_entry__.build() whenever /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("B2C_REFERENCE_SRC", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
# the files SURVEY.md §8(a) cites for the path, plus the consumers the parity tests replay
FILES = ["utils/embedder.py", "utils/image_features.py", "utils/nn_model.py", "_1_embed_with_CLIP.py", "_2_remove_duplicates.py",
         "_5_predict_labels.py"]


def make_ref(verbose: bool = True) -> str | None:
    if not os.path.isfile(os.path.join(SRC, "utils", "embedder.py")):
        return None  # not in the build container: keep whatever baseline/_ref already holds
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest, "note": "verbatim copies; see oracle/make_ref.py"}, fh, indent=1)
    if verbose:
        print(f"baseline/_ref: {len(manifest)} reference files copied from {SRC}")
    return DST


def verify_ref() -> bool:
    """True when baseline/_ref holds exactly the files its manifest lists (unmodified)."""
    mpath = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    files = json.load(open(mpath))["files"]
    for rel, sha in files.items():
        p = os.path.join(DST, rel)
        if not os.path.isfile(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != sha:
            return False
    return set(files) == set(FILES)


if __name__ == "__main__":
    print(make_ref() or "reference tree not present; nothing copied")
