"""Copies the UNMODIFIED reference (Fool-Yang/AlphaSnake-Zero, /root/reference/code/utils/*.py) to baseline/_ref/code/utils so that
bench.py can time the reference's own Python on the GPU box's host cores (`/root/reference` does not exist there).

baseline/_ref/ is git-ignored (never part of this repository's history) but travels with the repository snapshot to the GPU box.
`pip install /root/reference` is not applicable: the reference has no setup.py / pyproject.toml; this copy
is that install step.  Run by __graft_entry__.build() whenever /root/reference is present.

  python baseline/vendor_reference.py [source_dir]
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "code", "utils")
FILES = ("__init__.py", "game.py", "agent.py", "mp_game_runner.py", "pit_agent.py", "pit_mp_game_runner.py")


def vendor(src_root="/root/reference/code"):
    src = os.path.join(src_root, "utils")
    if not os.path.isdir(src):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(DST, f)
        if os.path.exists(s) and not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    ok = vendor(sys.argv[1]) if len(sys.argv) > 1 else vendor()
    print("vendored the reference to %s" % DST if ok else "reference not found; nothing copied")
