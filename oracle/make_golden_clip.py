"""Generate tests/golden/clip_patch.npz by executing the UNMODIFIED source of the reference's clip_patch
(FDLNet-master/utils/image_utils.py). The module itself imports skimage (absent here), so the function's source
lines are compiled on their own with torch in scope. Run in the build container only."""
from __future__ import annotations

import ast
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import clip_oracle  # noqa: E402

SRC = Path("/root/reference/FDLNet-master/utils/image_utils.py")


def reference_clip_patch():
    text = SRC.read_text()
    tree = ast.parse(text)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "clip_patch")
    code = compile(ast.Module(body=[fn], type_ignores=[]), str(SRC), "exec")
    ns = {"torch": torch}
    exec(code, ns)
    return ns["clip_patch"]


def main():
    ref = reference_clip_patch()
    byxc, scale, ori, im_info, images = clip_oracle.make_clip_inputs()
    out = ref(byxc, scale, ori, im_info, images, 32)
    out_noori = ref(byxc, scale, None, im_info, images, 32)
    np.savez_compressed(REPO / "tests" / "golden" / "clip_patch.npz", patches=out.numpy().astype(np.float32),
                        patches_no_ori=out_noori.numpy().astype(np.float32),
                        images_sum=np.float64(images.double().sum().item()))
    mine = clip_oracle.clip_patch(byxc, scale, ori, im_info, images, 32)
    print("reference vs oracle max abs:", (mine - out).abs().max().item(), out.shape)


if __name__ == "__main__":
    main()
