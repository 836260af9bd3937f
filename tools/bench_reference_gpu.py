#!/usr/bin/env python
"""Time the reference's OWN GPU paths on the same B200 as ours — a comparison arm,
never a product path.

    python tools/bench_reference_gpu.py [--reference baseline/_ref] [--model L --height 1024 --width 1216]
                                        [--batches 1,8] [--compile] [-o profiles/r2_reference_gpu.json]

Runs the UNMODIFIED reference module (``s2m2.core.model.s2m2.S2M2`` from ``baseline/_ref/src`` — staged by
``tools/stage_reference.py``; ``/root/reference`` does not exist on the GPU box) with

  * ``torch.manual_seed(0)`` + the reference constructor's default init (BASELINE.md §4 weights),
  * eager ``torch.amp.autocast(fp16)`` under ``torch.inference_mode()`` — exactly what
    ``run_stereo_matching`` does (model_utils.py:75-83), timed with the reference's own protocol (CUDA events around
    N forwards, ``torch.cuda.synchronize()`` after every forward) AND with the bench protocol (one sync at the end),
  * optionally ``torch.compile(model)`` (demo/visualize_2d_simple.py:36-37),
  * eager fp32 with TF32 disabled (the L/XL parity oracle) for the EPE of the fp16 path against it,

and our ``s2m2_b200.model.S2M2`` loaded with the SAME state-dict on the same inputs: pairs/s, and the EPE of both fp16
paths against the reference's fp32 output.  Prints one JSON object.
"""
import argparse
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, n, sync_each):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
        e1.record()
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--model", default="L")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1216)
    ap.add_argument("--batches", default="1,8")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--smooth", type=int, default=0, help="1: box-blurred inputs (tests' default), 0: BASELINE.md §4 raw noise")
    ap.add_argument("-o", "--out", default=None)
    a = ap.parse_args()

    sys.path.insert(0, os.path.join(a.reference, "src"))
    try:
        import open3d  # noqa: F401
    except Exception:
        sys.modules["open3d"] = types.ModuleType("open3d")
    from s2m2.core.model.s2m2 import S2M2 as RefS2M2
    from s2m2_b200.model import S2M2
    from s2m2_b200.weights import MODEL_TABLE, synthetic_pair

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True            # the demo sets it (visualize_2d_simple.py:17)
    dev = torch.device("cuda", 0)
    C, NTR = MODEL_TABLE[a.model]
    torch.manual_seed(0)
    ref = RefS2M2(feature_channels=C, dim_expansion=1, num_transformer=NTR, use_positivity=True, refine_iter=3).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev)
    ours = S2M2(C, 1, NTR, use_positivity=True, refine_iter=3, precision="fp16").eval()
    ours.my_load_state_dict(sd)
    ours = ours.to(dev)

    res = {"model": a.model, "height": a.height, "width": a.width, "weights": "torch.manual_seed(0) + reference ctor default init",
           "inputs": f"synthetic_pair(seed=1, smooth={bool(a.smooth)})", "torch": torch.__version__,
           "gpu": torch.cuda.get_device_name(0), "runs": []}
    for B in [int(b) for b in a.batches.split(",")]:
        left, right = synthetic_pair(B, a.height, a.width, 1, smooth=bool(a.smooth))
        left, right = left.to(dev), right.to(dev)
        row = {"batch": B}

        def eager16():
            with torch.inference_mode(), torch.amp.autocast(enabled=True, device_type="cuda", dtype=torch.float16):
                return ref(left, right)

        def ours16():
            with torch.inference_mode():
                return ours(left, right)

        try:
            # fp32 oracle (TF32 off)
            with torch.inference_mode():
                d32, o32, c32 = ref(left, right)
            torch.cuda.synchronize()
            for _ in range(3):
                eager16()
            ms_sync, (d16, o16, c16) = timed(eager16, a.steps, True)
            ms_nosync, _ = timed(eager16, a.steps, False)
            row["ref_eager_fp16"] = {"ms_per_forward_sync_each": ms_sync, "ms_per_forward": ms_nosync,
                                     "pairs_per_s": B * 1e3 / ms_nosync, "pairs_per_s_sync_each": B * 1e3 / ms_sync,
                                     "epe_vs_ref_fp32": (d16.float() - d32).abs().mean().item(),
                                     "linf_vs_ref_fp32": (d16.float() - d32).abs().max().item()}
            t0 = time.time()
            with torch.inference_mode():
                ms32, _ = timed(lambda: ref(left, right), 3, True)
            row["ref_eager_fp32_tf32off"] = {"ms_per_forward": ms32, "pairs_per_s": B * 1e3 / ms32}
        except torch.cuda.OutOfMemoryError as e:      # eager at batch 8 allocates every intermediate
            row["ref_eager_error"] = str(e)[:200]
            d32 = None
        ours.auto_graph = False
        for _ in range(3):
            ours16()
        ms_o_sync, (do, oo, co) = timed(ours16, a.steps, True)
        ms_o, _ = timed(ours16, a.steps, False)
        row["ours_fp16_eager_launches"] = {"ms_per_forward_sync_each": ms_o_sync, "ms_per_forward": ms_o, "pairs_per_s": B * 1e3 / ms_o,
                                           "pairs_per_s_sync_each": B * 1e3 / ms_o_sync}
        if d32 is not None:
            row["ours_fp16_eager_launches"].update({"epe_vs_ref_fp32": (do - d32).abs().mean().item(),
                                                    "linf_vs_ref_fp32": (do - d32).abs().max().item(),
                                                    "occ_linf": (oo - o32).abs().max().item(), "conf_linf": (co - c32).abs().max().item()})
        # the nn.Module itself replays a captured graph from the third same-shape call on (what run_stereo_matching's loop sees)
        ours.auto_graph = True
        for _ in range(3):
            ours16()
        ms_ag_sync, _ = timed(ours16, a.steps, True)
        row["ours_fp16_module_auto_graph"] = {"ms_per_forward_sync_each": ms_ag_sync, "pairs_per_s_sync_each": B * 1e3 / ms_ag_sync}
        ours._graphs = {}
        torch.cuda.empty_cache()
        from s2m2_b200.runtime import PipelinedRunner
        runner = PipelinedRunner(ours, B, a.height, a.width, dev, use_graph=True)
        runner.left.copy_(left)
        runner.right.copy_(right)
        for _ in range(3):
            runner.replay()
        ms_g, _ = timed(runner.replay, a.steps, False)
        ms_gs, _ = timed(runner.replay, a.steps, True)
        row["ours_fp16_graph"] = {"ms_per_forward": ms_g, "ms_per_forward_sync_each": ms_gs, "pairs_per_s": B * 1e3 / ms_g}
        del runner
        if a.compile:
            try:
                torch._dynamo.reset()
                cm = torch.compile(ref)

                def comp16():
                    with torch.inference_mode(), torch.amp.autocast(enabled=True, device_type="cuda", dtype=torch.float16):
                        return cm(left, right)
                t0 = time.time()
                comp16()
                torch.cuda.synchronize()
                t_compile = time.time() - t0
                for _ in range(2):
                    comp16()
                ms_c_sync, (dc, _, _) = timed(comp16, a.steps, True)
                ms_c, _ = timed(comp16, a.steps, False)
                row["ref_torch_compile_fp16"] = {"ms_per_forward_sync_each": ms_c_sync, "ms_per_forward": ms_c, "pairs_per_s": B * 1e3 / ms_c,
                                                 "compile_s": t_compile}
                if d32 is not None:
                    row["ref_torch_compile_fp16"]["epe_vs_ref_fp32"] = (dc.float() - d32).abs().mean().item()
            except Exception as e:  # noqa: BLE001 — the comparison arm must not take the job down
                row["ref_torch_compile_error"] = f"{type(e).__name__}: {str(e)[:300]}"
        res["runs"].append(row)
        print(json.dumps(row), flush=True)
        del left, right
        torch.cuda.empty_cache()
    text = json.dumps(res, indent=1)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
