"""Recipe that stages the UNMODIFIED reference sources the hot path needs into oracle/_ref/ (git-ignored, but
shipped to the GPU box by gpurun), so that the GPU-side tests and `bench.py --impl reference` can execute the
real reference code where /root/reference does not exist.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing is edited: files are copied byte for byte and
a manifest with their sha256 is written next to them.  Run in the build container (where /root/reference exists):

    python oracle/make_ref.py          (also called by __graft_entry__.build())

Copied: models/*.py, models/archs/*.py, models/losses/*.py and configs/*.yml — the files SURVEY.md §8 cites.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
PATTERNS = (("models", ".py"), ("models/archs", ".py"), ("models/losses", ".py"), ("configs", ".yml"))


def stage(src=SRC, dst=DST):
    """-> number of files staged (0 when the reference tree is absent: the GPU box uses what was shipped)"""
    if not os.path.isdir(src):
        return 0
    manifest = {}
    for rel, ext in PATTERNS:
        sdir = os.path.join(src, rel)
        ddir = os.path.join(dst, rel)
        os.makedirs(ddir, exist_ok=True)
        for name in sorted(os.listdir(sdir)):
            if not name.endswith(ext):
                continue
            s, d = os.path.join(sdir, name), os.path.join(ddir, name)
            shutil.copyfile(s, d)
            with open(d, "rb") as f:
                manifest[os.path.join(rel, name)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=src, files=manifest), f, indent=1, sort_keys=True)
    return len(manifest)


if __name__ == "__main__":
    n = stage()
    print(f"staged {n} reference files into {DST}" if n else f"{SRC} absent: nothing staged")
    sys.exit(0)
