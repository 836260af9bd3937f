#!/usr/bin/env python
"""Run the reference's UNMODIFIED demo/visualize_2d_simple.py with the B200-native S2M2 substituted.

    python tools/run_reference_demo.py --reference /path/to/s2m2 [--scratch DIR] [-- demo args...]

What it does (nothing in the reference tree is edited):
  * puts <reference>/src on sys.path and replaces the class at ``s2m2.core.model.s2m2.S2M2`` with
    ``s2m2_b200.model.S2M2`` *before* ``s2m2.core.utils.model_utils`` binds it
    (``from ..model.s2m2 import S2M2 as Model``, model_utils.py:6);
  * provides an ``open3d`` stub if the package is absent (model_utils.py:5 imports it unconditionally);
  * makes ``cv2.namedWindow/imshow/waitKey/destroyAllWindows`` no-ops when OpenCV is the headless build;
  * the demo derives its project root from its own location (visualize_2d_simple.py:15) and the reference tree
    may be read-only, so the demo file and the sample images are staged byte-for-byte into --scratch together with
    ``weights/pretrain_weights/CH{C}NTR{N}.pth`` (synthetic random-init weights are written if no checkpoint is
    given with --weights; there is no network for the HuggingFace checkpoints).
"""
import argparse
import os
import runpy
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_substitution(reference_root: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(reference_root, "src"))
    try:
        import open3d  # noqa: F401
    except Exception:
        sys.modules["open3d"] = types.ModuleType("open3d")
    import s2m2.core.model.s2m2 as ref_mod
    from s2m2_b200.model import S2M2
    ref_mod.S2M2 = S2M2
    import cv2
    try:
        cv2.namedWindow("__probe__")
        cv2.destroyAllWindows()
    except Exception:
        for name in ("namedWindow", "imshow", "destroyAllWindows", "resizeWindow", "moveWindow"):
            setattr(cv2, name, lambda *a, **k: None)
        cv2.waitKey = lambda *a, **k: -1
    return S2M2


def stage(reference_root: str, scratch: str, model_type: str, weights: str = None):
    from s2m2_b200.weights import MODEL_TABLE, random_state_dict
    import torch
    os.makedirs(os.path.join(scratch, "demo"), exist_ok=True)
    shutil.copy2(os.path.join(reference_root, "demo", "visualize_2d_simple.py"), os.path.join(scratch, "demo"))
    dst = os.path.join(scratch, "data", "samples", "Web")
    os.makedirs(dst, exist_ok=True)
    src = os.path.join(reference_root, "data", "samples", "Web")
    for f in os.listdir(src):
        shutil.copy2(os.path.join(src, f), dst)
    C, NTR = MODEL_TABLE[model_type]
    wdir = os.path.join(scratch, "weights", "pretrain_weights")
    os.makedirs(wdir, exist_ok=True)
    wpath = os.path.join(wdir, f"CH{C}NTR{NTR}.pth")
    if weights:
        shutil.copy2(weights, wpath)
    elif not os.path.exists(wpath):
        torch.save({"state_dict": dict(random_state_dict(C, NTR, 0))}, wpath)
    return os.path.join(scratch, "demo", "visualize_2d_simple.py")


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--scratch", default="/tmp/s2m2_b200_demo")
    ap.add_argument("--weights", default=None, help="a reference checkpoint CH{C}NTR{N}.pth (optional)")
    args, demo_args = ap.parse_known_args(argv)
    if demo_args and demo_args[0] == "--":
        demo_args = demo_args[1:]
    model_type = "S"
    if "--model_type" in demo_args:
        model_type = demo_args[demo_args.index("--model_type") + 1]
    install_substitution(args.reference)
    demo = stage(args.reference, args.scratch, model_type, args.weights)
    sys.argv = [demo] + demo_args
    runpy.run_path(demo, run_name="__main__")


if __name__ == "__main__":
    main()
