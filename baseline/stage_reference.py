"""Stages the UNMODIFIED reference (ngraymon/Pibronic, pure Python) into ``baseline/_ref`` so that
``bench.py --impl reference`` can time the reference's own ``block_compute_pm`` on the GPU box.

    python baseline/stage_reference.py [/root/reference]

``baseline/_ref`` is git-ignored (no reference source enters the history) but travels with the tree to
the GPU box.  The reference needs no build: the recipe copies the ``pibronic`` package as it lies under
the reference tree (``pip install --target`` would additionally want ``parse==1.8.2`` and the Julia bridge,
which the hot path never touches: SURVEY.md App. C).  A manifest with the SHA-256 of every file copied is
written next to it; ``baseline/ref_runner.py`` imports the staged package with stand-ins for the absent
third-party modules.
"""
import hashlib
import json
import os
import shutil
import sys
from os.path import abspath, dirname, isdir, join

HERE = dirname(abspath(__file__))
DEST = join(HERE, "_ref")


def stage(source="/root/reference", force=False):
    """copies <source>/pibronic -> baseline/_ref/pibronic; returns the destination or None if there is no source"""
    pkg = join(source, "pibronic")
    if not isdir(pkg):
        return None
    target = join(DEST, "pibronic")
    if isdir(target) and not force:
        return DEST
    if isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(pkg, target, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(target):
        for name in sorted(files):
            path = join(root, name)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, DEST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": source, "files": manifest}, fh, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    out = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", force=True)
    print(out if out else "no reference tree found: nothing staged")
