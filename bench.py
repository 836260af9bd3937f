#!/usr/bin/env python
"""bench.py — stereo pairs/s of the S2M2 forward on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--model L] [--height 1024] [--width 1216] [--batch B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)
  python bench.py --impl reference ...     (the reference algorithm's CPU path = the oracle port, host cores)

A "step" is one forward over `batch` synthetic stereo pairs per GPU (random-init weights of the named
architecture, synthetic images — there is no network for checkpoints/datasets).  Default batch = 8 pairs per GPU
(BASELINE config 5: batch 64 sharded over 8 GPUs, weak scaling keeps 8 per GPU at every N); the single-pair latency
(config 3, the protocol of the reference's README FPS table) is reported beside it as `single_pair`.
  value      pairs/s with the inputs already resident in HBM (CUDA-graph replay of the whole forward),
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e        the same metric through the public runtime call (s2m2_b200.runtime.PipelinedRunner.run) with HOST (pinned)
             inputs: H2D copy of both images and D2H read of disparity/occlusion/confidence inside the timed region, every
             step; the copies of neighbouring steps overlap the compute of the current one on a second stream.
  e2e_cabi   (N=1) the same step through the handle-level C-ABI (s2m2b200_forward: C++ schedule, eager launches, host buffers in and
             out, one synchronize per step) — what a host that is not Python gets.
  roofline   dominant kernel = the implicit-GEMM convolution; the heaviest conv shape of the model is timed
             alone with CUDA events; achieved = 2*M*N*K / time against MEASURED_PEAKS.json (burst figure); forward_frac =
             the whole forward (pairs/s x TFLOP/pair) against the sustained figure.
  cpu_baseline  the oracle (port of the reference's fp32 algorithm) on the host cores, rank 0, N=1 only: full-size pairs.
  --impl reference  the same CPU arm as its own JSON line: full-size pairs, one per forward, as many as fit S2M2_REF_BUDGET_S.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_PAIR = {  # 2*MAC over conv/linear/bmm, refine_iter=3  [TFLOP]
    ("S", 480, 640): 0.763, ("L", 1024, 1216): 14.451, ("XL", 2048, 2432): 133.535,
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "src": "measured"}
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_threads():
    """Threads for the CPU arm: the cgroup CPU quota if there is one (nproc can report far more cores than the
    container may use, and eager PyTorch over-subscribes badly), capped at 32."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, min(n, 32))


def time_oracle(sd, NTR, H, W, seed, budget_s, max_forwards):
    """Time the oracle port (fp32 eager torch ops on the host cores) on FULL-SIZE pairs of the named resolution: one pair per
    forward, as many forwards as fit the budget (at least one).  No crop, no extrapolation: pairs/s = forwards / seconds."""
    from oracle import s2m2_oracle as O
    from s2m2_b200.weights import synthetic_pair
    left, right = synthetic_pair(1, H, W, seed)
    wl, wr = left[..., :128, :160].contiguous(), right[..., :128, :160].contiguous()
    O.forward(sd, wl, wr, NTR, True, 3)                     # thread-pool / allocator warm-up on a small crop (untimed)
    times = []
    t_start = time.time()
    while len(times) < max_forwards:
        t0 = time.time()
        O.forward(sd, left, right, NTR, True, 3)
        times.append(time.time() - t0)
        if time.time() - t_start + times[-1] > budget_s:     # the next one would overrun the budget
            break
    return times


def run_reference(args, cfg):
    """--impl reference: the reference's CPU algorithm (oracle port, fp32 eager torch ops) on the host cores, rank 0 only."""
    from s2m2_b200.weights import MODEL_TABLE, random_state_dict
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C, NTR = MODEL_TABLE[args.model]
    cores = host_threads()
    torch.set_num_threads(cores)
    sd = random_state_dict(C, NTR, 0)
    H, W = args.height, args.width
    budget = float(os.environ.get("S2M2_REF_BUDGET_S", "100"))
    times = time_oracle(sd, NTR, H, W, 1, budget, args.steps)
    dt = sum(times) / len(times)
    pairs_per_s = 1.0 / dt
    sample = (f"oracle port fp32 eager on {cores} host threads: {len(times)} timed forward(s) of ONE full-size {W}x{H} pair each "
              f"(no crop, no extrapolation; {args.steps} requested, budget {budget:.0f} s)")
    cfg = dict(cfg)
    cfg["reference_sample"] = {"pairs_per_forward": 1, "height": H, "width": W, "forwards_timed": len(times),
                               "seconds": [round(t, 3) for t in times], "extrapolated": False}
    line = {"impl": "reference", "metric": "stereo pairs/sec", "value": pairs_per_s, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "steps_timed": len(times), "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": pairs_per_s, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pairs_per_s, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def heaviest_conv(model, B, H, W):
    """The conv launch with the most FLOPs in the schedule: found by recording conv shapes of one forward."""
    rec = {"best": None, "total": 0.0}      # only the heaviest launch keeps its tensors alive
    ops = model.ops
    orig = ops.conv2d

    def spy(x, cw, act, out, res=None, x2=None, **kw):
        M = out.shape[0] * out.shape[1] * out.shape[2] // (4 if cw.shuffle else 1)
        fl = 2.0 * M * cw.cout * cw.kh * cw.kw * cw.cin
        rec["total"] += fl
        if rec["best"] is None or fl > rec["best"][0]:
            rec["best"] = (fl, x, cw, act, out, res, x2)
        return orig(x, cw, act, out, res, x2, **kw)
    ops.conv2d = spy
    return rec, (lambda: setattr(ops, "conv2d", orig))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="s2m2_b200", choices=["s2m2_b200", "reference"])
    ap.add_argument("--model", default="L", choices=["S", "M", "L", "XL"])
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1216)
    ap.add_argument("--batch", type=int, default=8,
                    help="stereo pairs per GPU per step (default 8 = BASELINE config 5's share: batch 64 over 8 GPUs)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cabi", action="store_true", help="skip the handle-level C-ABI end-to-end leg (e2e_cabi)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    H, W, B = args.height, args.width, args.batch
    cfg = {"workload": f"{args.model} (CH/NTR per model_utils.py:12-17) {W}x{H} refine_iter=3, {B} pair(s)/GPU/step",
           "global_batch": B * world, "parallelism": f"dp{world} (batch-sharded pairs, NCCL weight broadcast only)",
           "precision": args.precision, "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"}
    if args.impl == "reference":
        return run_reference(args, cfg)

    from s2m2_b200.model import S2M2
    from s2m2_b200.weights import MODEL_TABLE, random_state_dict, synthetic_pair
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    C, NTR = MODEL_TABLE[args.model]
    dist = None
    bcast_bytes = 0
    nccl_log = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        # evidence of the communicator: NCCL's own init log (the "rank r nranks N ... Init COMPLETE" lines).  It goes to a per-rank
        # file and its communicator lines are re-emitted on this rank's stderr after the broadcast, so that stdout stays the one JSON
        # line and ranks do not overwrite each other when stderr is a file (NCCL opens its log file truncating).  A quieter preset
        # (unset / VERSION / WARN) is raised to INFO; INFO / TRACE / a log file chosen by the caller are left alone.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        if "NCCL_DEBUG_FILE" not in os.environ:
            nccl_log = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"s2m2_bench_nccl_rank{rank}_{os.getpid()}.log")
            os.environ["NCCL_DEBUG_FILE"] = nccl_log
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=10))

    model = S2M2(C, 1, NTR, use_positivity=True, refine_iter=3, precision=args.precision).eval()
    model.auto_graph = False          # the runtime below owns the captured graphs
    if rank == 0:
        model.my_load_state_dict(random_state_dict(C, NTR, 0))
    model = model.to(dev)
    if world > 1:   # the only collective on the path: ONE broadcast of the packed weight blob over NVLink
        from s2m2_b200.parallel import broadcast_packed
        from s2m2_b200.ops import CudaOps
        model.ops = CudaOps()
        bcast_bytes = broadcast_packed(model, dev, 0)
        torch.cuda.synchronize()
        print(f"[rank {rank}/{world}] packed weights received: {bcast_bytes / 1e6:.1f} MB (one NCCL broadcast from rank 0)", file=sys.stderr, flush=True)
        if nccl_log and os.path.exists(nccl_log):
            import re
            with open(nccl_log, errors="replace") as f:
                keep = [ln.rstrip() for ln in f if re.search(r"Init COMPLETE|nranks|NVLS|NCCL version", ln)]
            for ln in keep[:12]:
                print(f"[rank {rank}/{world}] {ln}", file=sys.stderr, flush=True)
            try:
                os.remove(nccl_log)
            except OSError:
                pass

    left_h, right_h = synthetic_pair(B, H, W, 1 + rank)
    left_h, right_h = left_h.pin_memory(), right_h.pin_memory()
    left, right = left_h.to(dev), right_h.to(dev)

    # ---- the serving runtime (s2m2_b200/runtime.py): warm-up, CUDA-graph capture of the forward, double-buffered host pipeline
    from s2m2_b200.runtime import PipelinedRunner
    with torch.inference_mode():
        model(left, right)                          # first call: creates the op backend, packs the weights
    torch.cuda.synchronize()
    n0 = model.ops.launch_count()
    runner = PipelinedRunner(model, B, H, W, dev, use_graph=not args.no_graph)      # 2 eager forwards (+ 1 captured)
    launches_per_fwd = (model.ops.launch_count() - n0) // (2 if args.no_graph else 3)
    graph = runner.graph
    runner.left.copy_(left)
    runner.right.copy_(right)
    left, right = runner.left, runner.right
    with torch.inference_mode():
        step = runner.replay                      # device-resident inputs: one graph replay per step
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- e2e: the runtime's public call with HOST inputs / outputs: every step copies its two pinned host images to the
        # device and its three maps back; the copies of step i+1 / i-1 ride a second stream under the compute of step i
        host_out = [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(3)]
        comp = torch.cuda.current_stream()

        def e2e_run(n):
            runner.run([(left_h, right_h)] * n, host_out=[host_out] * n)
            comp.wait_stream(runner.copy_stream)      # the last device->host copy is inside the timed region
        e2e_run(2)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_run(args.steps)
        e3.record()
        torch.cuda.synchronize()
        ms_e2e = e2.elapsed_time(e3)
        out = runner.out

        # ---- the same end-to-end call fed with uint8 images (4x fewer H2D bytes, bit-identical outputs) through
        # the same pipelined runtime: PipelinedRunner(input_dtype=uint8) -> its own captured graph
        e2e_u8 = None
        if rank == 0:
            lu_h, ru_h = left_h.to(torch.uint8).pin_memory(), right_h.to(torch.uint8).pin_memory()
            ru8 = PipelinedRunner(model, B, H, W, dev, input_dtype=torch.uint8, use_graph=not args.no_graph)
            nu = max(2, min(args.steps, 10))

            def u8_run(n):
                ru8.run([(lu_h, ru_h)] * n, host_out=[host_out] * n)
                comp.wait_stream(ru8.copy_stream)
            u8_run(2)
            torch.cuda.synchronize()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            u8_run(nu)
            u1.record()
            torch.cuda.synchronize()
            e2e_u8 = {"value": B * nu / (u0.elapsed_time(u1) * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * 3 * H * W,
                      "d2h_bytes_per_step": 3 * B * H * W * 4, "steps": nu,
                      "note": "uint8 host images through PipelinedRunner(input_dtype=uint8).run (graph replay, overlapped copies), this rank only"}
            del ru8

        # ---- the same end-to-end step through the handle-level C-ABI (s2m2b200_forward, host side in C++, eager launches — what a
        # host that is not Python calls): per step pinned host images -> device, the forward, the three maps -> pinned host memory,
        # one stream synchronize (no overlap between steps)
        e2e_cabi = None
        if rank == 0 and world == 1 and not args.no_cabi:
            from s2m2_b200.host import HostForward
            hf = HostForward(C, NTR, use_positivity=True, refine_iter=3, precision=args.precision)
            hf.load_state_dict(model.state_dict())
            dl, dr = torch.empty_like(left), torch.empty_like(right)
            nc = max(2, min(args.steps, 5))

            def cabi_run(n):
                for _ in range(n):
                    dl.copy_(left_h, non_blocking=True)
                    dr.copy_(right_h, non_blocking=True)
                    o = hf(dl, dr)
                    for k in range(3):
                        host_out[k].copy_(o[k], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            cabi_run(2)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = hf.lib.s2m2b200_launch_count()
            c0.record()
            cabi_run(nc)
            c1.record()
            torch.cuda.synchronize()
            e2e_cabi = {"value": B * nc / (c0.elapsed_time(c1) * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * 3 * H * W * 4,
                        "d2h_bytes_per_step": 3 * B * H * W * 4, "steps": nc, "gpu_launches": int(hf.lib.s2m2b200_launch_count() - n0),
                        "workspace_bytes": hf.workspace_bytes(B, H, W),
                        "note": "s2m2b200_create/set_param/finalize_params/workspace_bytes/forward (include/s2m2b200.h): C++ schedule, eager "
                                "launches, synchronous steps, this rank only"}
            del hf, dl, dr
            torch.cuda.empty_cache()

        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        per_rank = [B * args.steps / (ms * 1e-3)]
        print(f"[rank {rank}/{world}] {per_rank[0]:.2f} pairs/s device-resident, {B * args.steps / (ms_e2e * 1e-3):.2f} pairs/s end to end "
              f"({args.steps} steps of {B} pairs)", file=sys.stderr, flush=True)
        if dist:
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [B * args.steps / (float(a[0]) * 1e-3) for a in allt]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

        # ---- single-pair latency (BASELINE config 3, the protocol behind the reference's README FPS): batch 1, rank 0
        single = None
        if rank == 0 and B != 1:
            r1 = PipelinedRunner(model, 1, H, W, dev, use_graph=not args.no_graph)
            r1.left.copy_(left[:1])
            r1.right.copy_(right[:1])
            for _ in range(3):
                r1.replay()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n1 = max(args.steps, 5)
            torch.cuda.synchronize()
            a0.record()
            for _ in range(n1):
                r1.replay()
            a1.record()
            torch.cuda.synchronize()
            ms1 = a0.elapsed_time(a1) / n1
            single = {"batch": 1, "ms_per_pair": ms1, "pairs_per_s": 1e3 / ms1, "steps": n1}
            del r1

        # ---- roofline of the dominant kernel (implicit-GEMM conv), heaviest shape timed alone (rank 0)
        roof = None
        if rank == 0:
            rec, restore = heaviest_conv(model, B, H, W)
            model(left, right)
            restore()
            torch.cuda.synchronize()
            flops, x, cw, act, o, res, x2 = rec["best"]
            total_conv_flops = rec["total"]
            for _ in range(3):
                model.ops.conv2d(x, cw, act, o, None, x2)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            torch.cuda.synchronize()
            k0.record()
            for _ in range(reps):
                model.ops.conv2d(x, cw, act, o, None, x2)
            k1.record()
            torch.cuda.synchronize()
            kms = k0.elapsed_time(k1) / reps
            pk = _peaks()
            ach = flops / (kms * 1e-3) / 1e12
            traffic = None   # dram read+write bytes per launch of this kernel, from the committed ncu --set full capture
            tpath = os.path.join(ROOT, "profiles", "r1_conv_tc_traffic.json")
            Mrows = o.shape[0] * o.shape[1] * o.shape[2]
            if os.path.exists(tpath):
                for kname, kv in json.load(open(tpath))["kernels"].items():
                    if f"M={Mrows} N={cw.cout} K={cw.kh * cw.kw * cw.cin}" in kname:
                        traffic = kv["dram_read_bytes"] + kv["dram_write_bytes"]
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tensor_burst"], "unit": "TFLOP/s", "frac": ach / pk["tensor_burst"],
                    "traffic": traffic, "traffic_note": "dram__bytes_read.sum+dram__bytes_write.sum per launch, profiles/r1_conv_tc_ncu.md (single-CTA, batch 1) / r1_conv_pair_ncu.md (CTA pair, batch 8)",
                    "peak_source": pk["src"] + " bf16 burst (kernel timed alone)",
                    "kernel": f"implicit-GEMM conv {cw.name}: M={o.shape[0] * o.shape[1] * o.shape[2]} N={cw.cout} K={cw.kh * cw.kw * cw.cin}",
                    "kernel_ms": kms, "kernel_gflop": flops / 1e9, "conv_gflop_per_step": total_conv_flops / 1e9}

    pairs = B * world * args.steps
    value = pairs / (ms * 1e-3)
    if dist:
        dist.barrier()
        dist.destroy_process_group()      # before the JSON line: it stays the last line of stdout whatever NCCL logs at teardown
        dist = None
    if rank != 0:
        return
    tf = FLOP_PER_PAIR.get((args.model, H, W))
    pk = _peaks()
    if single:
        cfg["single_pair"] = single      # BASELINE config 3 (the README FPS protocol), measured in the same run on rank 0
    headline_cfg = (args.model, H, W) == ("L", 1024, 1216)
    vs_base = None
    if headline_cfg:
        vs_base = (single["pairs_per_s"] if single else (value / world if B == 1 else None))
        vs_base = vs_base / 11.2 if vs_base else None
    if roof is not None and tf:
        ach = value / world * tf
        roof["forward_achieved"] = ach
        roof["forward_peak"] = pk["tensor_sustained"]
        roof["forward_frac"] = ach / pk["tensor_sustained"]
        roof["forward_note"] = (f"whole forward: pairs/s per GPU x {tf} TFLOP/pair (SURVEY 8(d)) against the {pk['src']} bf16 SUSTAINED "
                                "peak; `frac` above is the heaviest conv launch timed alone against the burst peak")
    line = {"metric": "stereo pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": vs_base,
            "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * 3 * H * W * 4,
                    "d2h_bytes_per_step": 3 * B * H * W * 4,
                    "note": "s2m2_b200.runtime.PipelinedRunner.run: per step 2 pinned host images -> device, graph replay, 3 maps -> pinned "
                            "host; the copies of steps i+1 / i-1 overlap the compute of step i on a second stream (separately timed region)"},
            "gpu_launches": launches_per_fwd * args.steps, "launches_per_forward": launches_per_fwd,
            "cuda_graph": graph is not None, "clocks": clocks, "roofline": roof,
            "per_rank_pairs_per_s": per_rank, "weight_broadcast_bytes": bcast_bytes}
    if single:
        line["single_pair"] = single
    if e2e_u8:
        line["e2e_uint8"] = e2e_u8
    if e2e_cabi:
        line["e2e_cabi"] = e2e_cabi
    if tf:
        ach = value / world * tf
        line["roofline_forward"] = {"bound": "tensor", "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                                    "frac": ach / pk["tensor_sustained"], "tflop_per_pair": tf,
                                    "peak_source": pk["src"] + " bf16 sustained"}
    line["vs_baseline_note"] = ("single_pair.pairs_per_s / 11.2 FPS: the reference README's TensorRT fp16 number on ONE RTX 5090, "
                                "L@1216x1024, single-pair protocol (BASELINE.md section 1) - like for like; `value` is batched throughput")
    ref_gpu = os.path.join(ROOT, "profiles", "r2_reference_gpu.json")
    if headline_cfg and os.path.exists(ref_gpu):   # the reference's own GPU paths, measured on a B200 of this pool (tools/bench_reference_gpu.py)
        try:
            runs = {r["batch"]: r for r in json.load(open(ref_gpu))["runs"]}
            line["ref_gpu_eager_fp16"] = {f"batch{b}_pairs_per_s": runs[b]["ref_eager_fp16"]["pairs_per_s"] for b in runs if "ref_eager_fp16" in runs[b]}
            line["ref_gpu_compile"] = {f"batch{b}_pairs_per_s": runs[b]["ref_torch_compile_fp16"]["pairs_per_s"] for b in runs
                                       if "ref_torch_compile_fp16" in runs[b]}
            line["ref_gpu_note"] = "committed measurement (profiles/r2_reference_gpu.json), not re-timed in this run"
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        torch.set_num_threads(cores)
        times = time_oracle(random_state_dict(C, NTR, 0), NTR, H, W, 1, float(os.environ.get("S2M2_CPU_BASELINE_BUDGET_S", "30")), 2)
        dt = sum(times) / len(times)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": f"oracle port fp32 eager, {len(times)} forward(s) of one full-size {W}x{H} pair ({dt:.1f} s each), "
                                          "no crop / extrapolation"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
