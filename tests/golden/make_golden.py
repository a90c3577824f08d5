"""Generates tests/golden/*.npz from the one executable piece of the reference's dependency
stack available in the build container: python cv2 (4.13.0) `reprojectImageTo3D`, the function
WorldBase::insertDisparityImage calls at volumetric_map_base/src/world_base.cc:75.

Run from the repo root:  python tests/golden/make_golden.py
The GPU box has no need for cv2: tests only read the committed .npz files.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from volumetric_mapping_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_case(seed, rows=60, cols=80, downsample=1.0):
    rng = np.random.default_rng(seed)
    disp = rng.uniform(0.5, 128.0, size=(rows, cols)).astype(np.float32)
    sel = rng.uniform(size=(rows, cols))
    disp[sel < 0.10] = -1.0
    disp[(sel >= 0.10) & (sel < 0.12)] = 0.0
    Q = scenes.generate_q(Tx=-0.110074, left_cx=367.215 / 8, left_cy=248.375 / 8,
                          left_fx=458.654 / 8, left_fy=458.654 / 8, right_cx=366.0 / 8)
    if downsample != 1.0:
        # mimic world_base.cc:55-69 in numpy double (same IEEE ops) before handing Q to cv2
        f, f2 = downsample, downsample * downsample
        Q = Q.copy()
        Q[0, 0] /= f; Q[0, 3] /= f2; Q[1, 1] /= f; Q[1, 3] /= f2
        Q[2, 3] /= f2; Q[3, 2] /= f; Q[3, 3] /= f
    out = cv2.reprojectImageTo3D(disp, Q, handleMissingValues=True, ddepth=cv2.CV_32F)
    return disp, Q, out


def main():
    print("cv2", cv2.__version__)
    for name, seed, ds in (("reproject_small_a", 11, 1.0), ("reproject_small_b", 12, 2.0)):
        disp, Q, out = small_case(seed, downsample=ds)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), disparity=disp, Q=Q, xyz=out,
                            cv2_version=cv2.__version__)
    # No-missing-value image: the image minimum is a legitimate disparity and still maps to 10000
    rng = np.random.default_rng(13)
    disp = rng.uniform(1.0, 64.0, size=(24, 32)).astype(np.float32)
    Q = scenes.generate_q(-0.2, 16.0, 12.0, 40.0, 40.0, 16.0)   # Q[3,3] == 0
    out = cv2.reprojectImageTo3D(disp, Q, handleMissingValues=True, ddepth=cv2.CV_32F)
    np.savez_compressed(os.path.join(HERE, "reproject_nomissing.npz"), disparity=disp, Q=Q,
                        xyz=out, cv2_version=cv2.__version__)
    # cfg2 at full size: commit only a checksum of cv2's output bytes
    c = scenes.cfg2_disparity()
    out = cv2.reprojectImageTo3D(c["disparity"], c["Q"], handleMissingValues=True,
                                 ddepth=cv2.CV_32F)
    digest = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
    valid = int(np.sum((out[..., 2] != 10000.0) & ~np.isinf(out[..., 2]) & ~(out[..., 2] < 0)))
    np.savez_compressed(os.path.join(HERE, "reproject_cfg2_digest.npz"), sha256=digest,
                        valid_pixels=valid, cv2_version=cv2.__version__)
    print("cfg2 digest", digest, "valid", valid)


if __name__ == "__main__":
    main()
