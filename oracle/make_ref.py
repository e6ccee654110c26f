"""TEST / BASELINE INFRASTRUCTURE — recipe that stages the UNMODIFIED reference modules of the ISTVT path under
oracle/_ref/ so that bench.py's `--impl reference` arm and `cpu_baseline` leg can time the reference itself (kind
"reference") on the GPU box, where /root/reference does not exist.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference is Python: there is nothing to compile, the "build output" is the set of source files the path imports —
found by loading `network.vivit.vivit` through oracle/reference_shim.py and listing the `network.*` modules that ended
up in sys.modules (19 files, ~200 KB).  They are copied byte for byte, with their relative paths, into oracle/_ref/,
which is git-ignored (no reference source enters the history) but not gpurun-ignored (it travels to the GPU box).
Nothing in the product package reads oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "oracle", "_ref")


def stage(reference_root: str = "/root/reference") -> int:
    if not os.path.isdir(os.path.join(reference_root, "network", "vivit")):
        print(f"make_ref: no reference checkout at {reference_root}; nothing staged")
        return 0
    sys.path.insert(0, ROOT)
    os.environ["ISTVT_REFERENCE_ROOT"] = reference_root
    from oracle import reference_shim as shim
    saved = {k: v for k, v in sys.modules.items() if k == "network" or k.startswith("network.")}
    shim._loaded = None
    shim.REFERENCE_ROOT = reference_root
    shim.load()
    files = sorted({os.path.relpath(m.__file__, reference_root) for k, m in sys.modules.items()
                    if (k == "network" or k.startswith("network.")) and getattr(m, "__file__", None)
                    and os.path.abspath(m.__file__).startswith(os.path.abspath(reference_root))})
    for k in [k for k in sys.modules if k == "network" or k.startswith("network.")]:
        del sys.modules[k]
    sys.modules.update(saved)
    shim._loaded = None
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in files:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference_root, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "files": manifest}, f, indent=1)
    print(f"make_ref: staged {len(files)} unmodified reference files under oracle/_ref/")
    return len(files)


if __name__ == "__main__":
    stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
