"""Recipe for ``oracle/_ref`` (TEST INFRASTRUCTURE): the reference's own processor module, taken UNMODIFIED from where it
lies under /root/reference, so that the GPU box — which has no /root/reference — can run the reference's own
``IPAttnProcessor2_0`` / ``AttnProcessor2_0`` classes as the checker (``tests/``, ``__graft_entry__.smoke()``) and as the
CPU arm of ``bench.py`` (``cpu_baseline`` / ``--impl reference``).

``APadapter/ap_adapter/attention_processor.py`` imports only torch ("What's missing" item 9), so it needs no
build: this script copies that one file into ``oracle/_ref/`` and records its sha256.  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box like the built ``libapx.so``.
The AudioMAE modules are not vendored: they import timm / librosa / matplotlib, which this image lacks — they stay pinned
through ``tests/golden/*.npz`` (``oracle/gen_golden.py``).

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("APX_REFERENCE_ROOT", "/root/reference")
FILES = {"attention_processor.py": os.path.join("APadapter", "ap_adapter", "attention_processor.py")}


def build(out_dir: str | None = None) -> dict | None:
    """Returns the manifest, or None when the reference tree is absent (GPU box: the prebuilt directory is used as is)."""
    out_dir = out_dir or os.path.join(HERE, "_ref")
    if not all(os.path.isfile(os.path.join(REF_ROOT, rel)) for rel in FILES.values()):
        return None
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"source_root": REF_ROOT, "files": {}}
    for name, rel in FILES.items():
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(out_dir, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest["files"][name] = {"from": rel, "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


if __name__ == "__main__":
    m = build()
    print("oracle/_ref:", "reference tree not present, nothing done" if m is None else json.dumps(m["files"]))
