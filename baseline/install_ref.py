#!/usr/bin/env python
"""Install the UNMODIFIED reference model code for bench.py's `--impl reference` arm and its `cpu_baseline` /
`torch_eager_same_gpu` legs.

The reference (augustgw/early-exit-transformer) is a tree of plain Python scripts: no setup.py / pyproject, so the base
contract's `pip install --target baseline/_ref /root/reference` has nothing to install.  What the hot path needs from it is
importable as-is: `models/` (Early_conformer, Splitformer, full_conformer and the blocks they pull in),
`util/model_utils.py` (initialize_weights) and `util/noam_opt.py` (NoamOpt).  This script copies exactly those files, byte for
byte, into `baseline/_ref/` -- git-ignored (never part of the repo's history), NOT gpurun-ignored (travels to the GPU box,
which has no /root/reference).  `train.py`, `inference.py`, `util/beam_infer.py`, `util/tokenizer.py` are NOT copied: they fail at
import in this image (flashlight-text / editdistance are missing).

Run in the authoring container:   python baseline/install_ref.py      (__graft_entry__.build() does it when /root/reference exists)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
FILES = ["util/__init__.py", "util/model_utils.py", "util/noam_opt.py"]


def install(src: str = SRC, dst: str = DST) -> bool:
    if not os.path.isdir(os.path.join(src, "models")):
        return False
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(os.path.join(src, "models"), os.path.join(dst, "models"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "#*#", "*~"))
    for rel in FILES:
        p = os.path.join(src, rel)
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(dst, rel))
        elif rel.endswith("__init__.py"):
            open(os.path.join(dst, rel), "w").close()
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as f:
        f.write(src + "\n")
    return True


if __name__ == "__main__":
    ok = install()
    print("baseline/_ref installed" if ok else f"{SRC} not present: nothing installed")
    sys.exit(0 if ok else 1)
