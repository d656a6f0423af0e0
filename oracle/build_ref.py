"""Recipe for ``oracle/_ref``: the part of the UNMODIFIED reference that implements the hot path, copied verbatim from
``/root/reference`` so that the CPU arm of ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) can time the
reference itself on the GPU box, where ``/root/reference`` does not exist.

    python oracle/build_ref.py          # build container only; __graft_entry__.build() runs it when /root/reference exists

``oracle/_ref`` is a build output (git-ignored, like the compiled library; it travels with the snapshot): nothing of the
reference is committed to this repository. Copied: ``koafusion/__init__.py``, ``koafusion/models/*.py`` (pure PyTorch +
torchvision + einops) and ``koafusion/various/_losses.py`` (FocalLoss). Test infrastructure only: the product
never imports it (tests/test_abi.py guards that).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ["koafusion/__init__.py", "koafusion/various/_losses.py"]


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "koafusion", "models")):
        print(f"{SRC} not present: oracle/_ref left as it is")
        return 0
    files = list(FILES) + sorted(os.path.join("koafusion/models", f) for f in os.listdir(os.path.join(SRC, "koafusion/models"))
                                 if f.endswith(".py"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, files=manifest), f, indent=1)
    print(f"oracle/_ref: {len(files)} files copied from {SRC}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
