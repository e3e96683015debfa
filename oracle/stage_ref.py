#!/usr/bin/env python
"""ORACLE (test infrastructure, NOT product code) -- stage the UNMODIFIED reference for the GPU box.

    python oracle/stage_ref.py          # /root/reference/research  ->  oracle/_ref/research  (byte-for-byte copy)

``oracle/_ref/`` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the copy travels
with the snapshot: ``bench.py --impl reference`` and the ``cpu_baseline`` / ``gpu_library_baseline`` legs then time the
reference's own ``Learner.action_sample`` (research/finetune_omtm/learner.py:329-417) on the box, instead of the oracle
port.  ``__graft_entry__.build()`` runs this whenever /root/reference exists.  A manifest (file list + sha256 of every file,
source commit if known) is written next to the copy so a reader can check nothing was edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"


def stage(src_root: str = SRC_DEFAULT, quiet: bool = False) -> bool:
    src = os.path.join(src_root, "research")
    if not os.path.isdir(src):
        if not quiet:
            print(f"stage_ref: {src} not found; nothing staged")
        return False
    dst = os.path.join(DST, "research")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.png", "*.gif", "*.mp4"))
    manifest = {}
    for dirpath, _, files in os.walk(dst):
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump({"source": src_root, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=0, sort_keys=True)
    if not quiet:
        print(f"stage_ref: {len(manifest)} files -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT) else 1)
