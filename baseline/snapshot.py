"""BASELINE / HARNESS INFRASTRUCTURE ONLY -- snapshot of the reference's network definitions.

BASELINE.json configs 4 and 5 (PoseGenerator / FaceGenerator) use the reference's OWN generator code,
unchanged, as the harness around the warping ops ("stock torch.nn, used unchanged
as the harness").  /root/reference does not exist on the GPU box, so `snapshot()` -- called by
`__graft_entry__.build()` in the build container -- copies the four network-definition files into the
git-ignored `baseline/_ref/` (it travels with the gpurun snapshot; nothing of it is ever committed).
The copies are byte-identical; `compat.install(reference_root=root())` puts them on `model.networks`.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["model/networks/generator.py", "model/networks/base_function.py", "model/networks/base_network.py",
         "model/networks/external_function.py"]


def snapshot(reference: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(reference, "model", "networks")):
        return False
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, rel), dst)
    return True


def root() -> str | None:
    """directory to hand to compat.install(reference_root=...), or None when no snapshot travelled"""
    return DEST if all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES) else None
