"""Golden vectors for the 8-point fallback of polynomial_triangulation (reference
lib/utils/triangulation.py:213-217), produced with the UNMODIFIED reference functions and the
installed OpenCV.  Build container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_8point.py

  f8 / corrected / x_8pt : the fallback branch on its own -- cv2.findFundamentalMat(FM_8POINT),
      cv2.correctMatches with it, the reference's linear_eigen_triangulation -- on the noisy
      ring-camera pairs of tests/golden_inputs.py::triangulation_case;
  x_same / st_same       : the reference's polynomial_triangulation with P2 == P1 (F = 0, the
      correction is NaN for every joint, so the reference itself takes the fallback)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402

tri = refshim.ref().triangulation
u1, u2, P1, P2, X = gi.triangulation_case()
f8, c1, c2, x8, xs, ss = [], [], [], [], [], []
for i in range(len(u1)):
    F = cv2.findFundamentalMat(u1[i], u2[i], cv2.FM_8POINT)[0]
    a, b = cv2.correctMatches(F, u1[i].reshape(1, -1, 2), u2[i].reshape(1, -1, 2))
    x, _ = tri.linear_eigen_triangulation(a[0], P1[i], b[0], P2[i])
    f8.append(F); c1.append(a[0]); c2.append(b[0]); x8.append(x)
    x, st = tri.polynomial_triangulation(u1[i], P1[i], u2[i], P1[i])
    xs.append(x); ss.append(np.asarray(st).astype(np.int64))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "triangulation_8point.npz"),
                    f8=np.asarray(f8), corrected_u1=np.asarray(c1), corrected_u2=np.asarray(c2),
                    x_8pt=np.asarray(x8), x_same=np.asarray(xs), st_same=np.asarray(ss))
print("wrote triangulation_8point", np.asarray(x8).shape, np.asarray(xs)[0, :2], np.asarray(ss)[0, :3])
