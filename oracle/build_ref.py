"""Recipe for ``oracle/_ref`` (TEST INFRASTRUCTURE): the reference's own processor module, COMPILED from where it lies under
/root/reference, so that the GPU box — which has no /root/reference — can run the reference's own ``IPAttnProcessor2_0`` /
``AttnProcessor2_0`` classes as the checker (``tests/``, ``__graft_entry__.smoke()``) and as the CPU arm of ``bench.py``
(``cpu_baseline`` / ``--impl reference``).

``APadapter/ap_adapter/attention_processor.py`` imports only torch ("What's missing" item 9), so it builds from its
one source file: this script byte-compiles it (``py_compile``, CPython's own compiler — nothing of ours in between) into
``oracle/_ref/attention_processor.pyc`` and records the sha256 of the source it was compiled from, of the output, and the
interpreter's bytecode magic.  No reference SOURCE is copied anywhere: ``oracle/_ref/`` holds build outputs only, is git-ignored
(nothing of the reference enters the history) but NOT gpurun-ignored, so it travels to the GPU box like the built ``libapx.so``.
The AudioMAE modules are not built: they import timm / librosa / matplotlib, which this image lacks — they stay pinned through
``tests/golden/*.npz`` (``oracle/gen_golden.py``).

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import py_compile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("APX_REFERENCE_ROOT", "/root/reference")
FILES = {"attention_processor.pyc": os.path.join("APadapter", "ap_adapter", "attention_processor.py")}


def _sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(out_dir: str | None = None) -> dict | None:
    """Returns the manifest, or None when the reference tree is absent (GPU box: the prebuilt directory is used as is)."""
    out_dir = out_dir or os.path.join(HERE, "_ref")
    if not all(os.path.isfile(os.path.join(REF_ROOT, rel)) for rel in FILES.values()):
        return None
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"source_root": REF_ROOT, "python_magic": importlib.util.MAGIC_NUMBER.hex(), "files": {}}
    for name, rel in FILES.items():
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(out_dir, name)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        manifest["files"][name] = {"from": rel, "source_sha256": _sha256(src), "sha256": _sha256(dst)}
        stale = os.path.join(out_dir, os.path.basename(rel))   # earlier form of this recipe placed the module itself here
        if os.path.isfile(stale):
            os.remove(stale)
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


if __name__ == "__main__":
    m = build()
    print("oracle/_ref:", "reference tree not present, nothing done" if m is None else json.dumps(m["files"]))
