#!/usr/bin/env python
"""TEST INFRASTRUCTURE — golden vectors for the rectification kernel (s2m2b200_remap_u8) made by OpenCV itself:
cv2.remap exactly as image_utils.py:120-134 calls it (uint8 image, float32 maps, INTER_LINEAR, constant border), on maps produced
the way the calibration tools produce them (cv2.stereoRectify + cv2.initUndistortRectifyMap, calib_utils.py) plus a synthetic map
that exercises the borders, exact-integer coordinates and the round-half-even ties of the 1/32-pixel quantisation.

    python oracle/make_golden_remap.py            # -> tests/golden/remap_u8.npz   (needs cv2; 4.13.0 here)
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rectification_maps(w, h, roll, pitch, yaw):
    """Maps of a slightly mis-aligned stereo rig, the same cv2 calls as calib_utils.compute_stereo_rectification."""
    K = np.array([[0.9 * w, 0, w / 2 - 3.5], [0, 0.9 * w, h / 2 + 2.25], [0, 0, 1]], np.float64)
    D = np.array([-0.05, 0.01, 0.0005, -0.0003, 0.0], np.float64)
    rx, ry, rz = np.deg2rad([roll, pitch, yaw])
    Rx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    Ry = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    Rz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    T = np.array([-0.12, 0.001, 0.0005])
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K, D, K, D, (w, h), R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
    lx, ly = cv2.initUndistortRectifyMap(K, D, R1, P1, (w, h), cv2.CV_32FC1)
    rx_, ry_ = cv2.initUndistortRectifyMap(K, D, R2, P2, (w, h), cv2.CV_32FC1)
    return lx, ly, rx_, ry_


def main():
    rng = np.random.default_rng(11)
    H, W = 96, 136
    left = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    right = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    out = {"left": left, "right": right}
    cands = [(0.0, 0.0, 0.0), (-1.5, 1.0, -0.7)]
    for i, (r, p, y) in enumerate(cands):
        lx, ly, rx, ry = rectification_maps(W, H, r, p, y)
        out[f"c{i}_lx"], out[f"c{i}_ly"], out[f"c{i}_rx"], out[f"c{i}_ry"] = lx, ly, rx, ry
        out[f"c{i}_left"] = cv2.remap(left, lx, ly, cv2.INTER_LINEAR, cv2.BORDER_CONSTANT)       # image_utils.py:120-126
        out[f"c{i}_right"] = cv2.remap(right, rx, ry, cv2.INTER_LINEAR, cv2.BORDER_CONSTANT)    # image_utils.py:128-134
    # synthetic map into a differently-sized destination: out-of-image taps on all four sides, integer / half / tie coordinates
    Hd, Wd = 72, 160
    yy, xx = np.meshgrid(np.arange(Hd, dtype=np.float32), np.arange(Wd, dtype=np.float32), indexing="ij")
    mx = (xx * 0.93 - 9.3 + 3 * np.sin(yy / 7)).astype(np.float32) + rng.normal(0, 0.3, (Hd, Wd)).astype(np.float32)
    my = (yy * 1.5 - 8.1 + 2 * np.cos(xx / 9)).astype(np.float32) + rng.normal(0, 0.3, (Hd, Wd)).astype(np.float32)
    mx[0, :40] = np.arange(40); my[0, :40] = 3
    mx[1, :40] = np.arange(40) + 0.5; my[1, :40] = 3.5
    mx[2, :40] = np.arange(40) + 1 / 64; my[2, :40] = 3 + 3 / 64
    mx[3, :40] = -1 + np.arange(40) * 0.25; my[3, :40] = -1.0
    mx[4, :40] = W - 2 + np.arange(40) * 0.125; my[4, :40] = H - 1.5
    out["syn_mx"], out["syn_my"] = mx, my
    out["syn_out"] = cv2.remap(left, mx, my, cv2.INTER_LINEAR, cv2.BORDER_CONSTANT)
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(ROOT, "tests", "golden", "remap_u8.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
