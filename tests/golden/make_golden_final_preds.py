"""Golden vectors for get_final_preds: the UNMODIFIED reference lib/core/inference.py:43-68
(through oracle/refshim.py; cv2.getAffineTransform from the installed OpenCV) on the seeded
heat-maps of tests/golden_inputs.py::final_preds_case.  Build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_final_preds.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

r = refshim.ref()
hm, center, scale = gi.final_preds_case()
out = {}
for pp in (True, False):
    cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(POST_PROCESS=pp))
    preds, maxvals = r.inference.get_final_preds(cfg, hm.copy(), center, scale)
    out["preds_pp%d" % pp] = preds
    out["maxvals_pp%d" % pp] = maxvals
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "final_preds.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
