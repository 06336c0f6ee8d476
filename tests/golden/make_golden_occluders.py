"""Golden vectors for the synthetic-occlusion augmentation: the UNMODIFIED reference
get_single_patch_sample (lib/utils/img_utils.py:246-298) with `occluder` set, i.e. including
occlude_with_objects / paste_over / resize_by_factor (lib/utils/augmentation.py:61-123), on the
seeded frames and occluders of tests/golden_inputs.py.  Build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_occluders.py"""
import importlib
import os
import random
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

refshim.ref()
iu = importlib.import_module("lib.utils.img_utils")
il = importlib.import_module("lib.core.integral_loss")
MEAN = np.array([123.675, 116.280, 103.530])
STD = np.array([58.395, 57.120, 57.375])
occluders = gi.occluder_set()
rec = {}
with tempfile.TemporaryDirectory() as tmp:
    for tag in gi.PATCH_CASES:
        img, box, joints, joints_vis, pw, ph, seed = gi.frame_case(tag)
        path = os.path.join(tmp, tag + ".png")
        assert cv2.imwrite(path, img)
        for aug in (False, True):
            np.random.seed(seed + 7); random.seed(seed + 7)
            patch, label, weight, s2, r2 = iu.get_single_patch_sample(
                path, box[0], box[1], box[2], box[3], joints.copy(), joints_vis.copy(), [], None, pw, ph,
                2000.0, 2000.0, MEAN, STD, aug, il.generate_joint_location_label, occluder=occluders)
            k = tag + ("_aug" if aug else "")
            rec[k + "_patch"] = patch
            rec[k + "_label"] = np.asarray(label, dtype=np.float64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "patch_occluders.npz"),
                    cv2_version=np.array(cv2.__version__), **rec)
print("wrote patch_occluders", {k: v.shape for k, v in rec.items() if k.endswith("_patch")})
