#!/usr/bin/env python
"""Recipe: stage the UNMODIFIED reference files of the hot path into oracle/_ref/ so that they travel to the GPU box.

    python oracle/stage_ref.py            (also run by __graft_entry__.build() wherever /root/reference exists)

The reference is pure Python, so "building" it means copying the few files of the path — byte for byte, same
directory layout — to where a box without /root/reference can import them:

    everglades-server/everglades_server/{server.py, definitions.py}     the game (server.py:211-279 game_turn, ...)
    gym-everglades/gym_everglades/{__init__.py, envs/*.py}              the gym wrapper (everglades_env.py:32-116)
    config/{DemoMap, UnitDefinitions, GameSetup}.json                   its game files

oracle/_ref/ is git-ignored (never part of the history: reference sources are not copied into the repo) but not
gpurun-ignored.  It is TEST INFRASTRUCTURE like the rest of oracle/: used by bench.py's `cpu_baseline` /
`--impl reference` legs to time the real reference on the box's host cores, and by nothing in the product.
A manifest with the SHA-256 of every staged file is written next to them.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("EVG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "everglades-server/everglades_server/server.py",
    "everglades-server/everglades_server/definitions.py",
    "gym-everglades/gym_everglades/__init__.py",
    "gym-everglades/gym_everglades/envs/__init__.py",
    "gym-everglades/gym_everglades/envs/everglades_env.py",
    "gym-everglades/gym_everglades/envs/everglades_renderer.py",
    "config/DemoMap.json",
    "config/UnitDefinitions.json",
    "config/GameSetup.json",
]


def stage(force: bool = False) -> bool:
    """Copy the files if the source tree is present. Returns whether oracle/_ref/ is complete afterwards."""
    if os.path.isdir(SRC):
        manifest = {}
        for rel in FILES:
            s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            data = open(s, "rb").read()
            if force or not os.path.isfile(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
            manifest[rel] = hashlib.sha256(data).hexdigest()
        with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
            json.dump({"source": SRC, "sha256": manifest}, f, indent=1, sort_keys=True)
    return all(os.path.isfile(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print("oracle/_ref/ %s" % ("staged" if ok else "INCOMPLETE (no %s here)" % SRC))
    sys.exit(0 if ok else 1)
