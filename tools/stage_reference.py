#!/usr/bin/env python
"""Stage the UNMODIFIED reference into ``baseline/_ref/`` so that it travels to the GPU box with gpurun.

    python tools/stage_reference.py [--reference /root/reference]

``baseline/_ref/`` is git-ignored (never part of the history, never imported by the product) and holds a verbatim copy of
the reference's ``src/``, ``demo/visualize_2d_simple.py`` and ``data/samples/Web``: it exists so that

  * ``tools/bench_reference_gpu.py`` can time the reference's own GPU paths (eager ``autocast(fp16)``,
    ``torch.compile``) on the same B200 as ours, and
  * ``tools/run_reference_demo.py --reference baseline/_ref`` can run the unmodified demo with the CUDA backend.

The reference is pure Python with ``torch, opencv-python, numpy`` as its only dependencies (pyproject.toml), all present in
this image; ``pip install --target`` would produce exactly the ``src/s2m2`` package copied here (recorded in DESIGN.md).
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(a.reference):
        raise SystemExit(f"{a.reference} not found (the GPU box only uses the copy already staged in baseline/_ref)")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(os.path.join(a.reference, "src"), os.path.join(dst, "src"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(dst, "demo"))
    shutil.copy2(os.path.join(a.reference, "demo", "visualize_2d_simple.py"), os.path.join(dst, "demo"))
    shutil.copytree(os.path.join(a.reference, "data", "samples", "Web"), os.path.join(dst, "data", "samples", "Web"))
    n = sum(len(f) for _, _, f in os.walk(dst))
    print(f"staged {n} files into {dst}")


if __name__ == "__main__":
    main()
