"""End-to-end parity of the CUDA path (through the nn.Module surface / C-ABI) against
  (a) golden vectors produced by the real reference (tests/golden, oracle/make_golden.py),
  (b) the oracle restatement run on the same seeded inputs (on the GPU in fp32 with TF32 disabled for the
      larger sizes so that it finishes in seconds),
  (c) size-independent properties at the BASELINE sizes.
Tolerances (north_star): EPE < 1e-3 px against the fp32 reference in BOTH precisions, and the integer contract — arg-max index
(submodules.py:226), conf > .2 (refinenet.py:65) and the per-iteration x - disp >= 0 mask (s2m2.py:179) — bit-identical in both
precisions at the BASELINE sizes, on the softened test inputs AND on the un-softened BASELINE.md §4 inputs.  The tiny golden models
(C = 32/64) in fp16 are additionally judged against the reference algorithm's own autocast-fp16 behaviour.
"""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import s2m2_oracle as O
from s2m2_b200.model import S2M2
from s2m2_b200.weights import random_state_dict, synthetic_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(C, NTR, sd, pos=True, it=3, precision="fp32"):
    m = S2M2(C, 1, NTR, use_positivity=pos, refine_iter=it, precision=precision).eval()
    m.my_load_state_dict(sd)
    return m.to(DEV)


def _oracle_gpu(sd, left, right, NTR, pos, it, taps=None):
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sdg = {k: v.to(DEV) for k, v in sd.items()}
        return O.forward(sdg, left.to(DEV), right.to(DEV), NTR, pos, it, taps)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _decision_mismatches(taps, ot, n_it):
    """Mismatch counts of the three discrete decisions of the path: prob_max_ind (submodules.py:226), conf > .2
    (refinenet.py:65) and the occlusion mask x - disp >= 0 of every refine iteration (s2m2.py:179; the refiner's sigmoid
    output is never exactly 0, so the zero pattern of occ after the iteration IS the mask)."""
    flips = int((taps["disp_init.argmax"].cpu() != ot["disp_init.argmax"].cpu()).sum())
    conf = int(((taps["disp_init.conf"].cpu() > 0.2) != (ot["disp_init.conf"].cpu() > 0.2)).sum())
    occ = [int(((taps[f"refiner.{it}.occ"].cpu() == 0) != (ot[f"refiner.{it}.occ"].cpu() == 0)).sum()) for it in range(n_it)]
    return flips, conf, occ


def _margin_report(taps_mine, ot):
    """arg-max flips of prob_max_ind (submodules.py:226) and the reference's own top-1/top-2 margin at those pixels."""
    prob = ot["disp_init.prob"]
    top2 = prob.topk(2, dim=3).values
    margin = ((top2[..., 0] - top2[..., 1]) / top2[..., 0].clamp(min=1e-30))
    flips = taps_mine["disp_init.argmax"] != ot["disp_init.argmax"]
    return flips, margin


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fp32_path_matches_reference_golden(name):
    cfg, g = load_golden(name)
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    m = _model(cfg["C"], cfg["NTR"], sd, cfg["pos"], cfg["it"], "fp32")
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    m.taps = {}
    d, o, c = m(left.to(DEV), right.to(DEV))
    torch.cuda.synchronize()
    d, o, c = d.cpu(), o.cpu(), c.cpu()
    assert d.shape == g["disp"].shape
    err = (d - g["disp"]).abs()
    epe = err.mean().item()
    print(f"{name}: EPE {epe:.3e} max {err.max().item():.3e}")
    # stages upstream of the hard decisions: always tight
    assert (m.taps["disp_init.cv"].cpu() - g["tap:disp_init#0.3"]).abs().max() < 5e-3
    if cfg["pos"]:
        assert epe < 1e-3
        assert err.max() < 2e-2
        assert (o - g["occ"]).abs().max() < 2e-3
        assert (c - g["conf"]).abs().max() < 2e-3
        assert (m.taps["disp_init.disp"].cpu() - g["tap:disp_init#0.0"]).abs().max() < 2e-3
    else:
        # use_positivity=False with random-init weights: the un-masked matching distribution has exact near-ties
        # (the reference's own autocast-fp16 flips ~2/3 of the arg-maxes here, tools/diag_fp16.py), so a flip of
        # prob_max_ind is legitimate iff the reference's top-1/top-2 margin at that pixel is at rounding level.
        ot = {}
        _oracle_gpu(sd, left, right, cfg["NTR"], cfg["pos"], cfg["it"], ot)
        flips, margin = _margin_report(m.taps, ot)
        print(f"   arg-max flips {int(flips.sum())} / {flips.numel()}, max reference margin at a flip "
              f"{float(margin[flips].max()) if flips.any() else 0.0:.2e}")
        assert flips.float().mean() < 0.02
        assert (not flips.any()) or float(margin[flips].max()) < 1e-4
        assert err.median() < 1e-4                 # everything not downstream of a flipped tie
        assert (err > 1e-2).float().mean() < 0.05


def test_fp32_argmax_and_masks_bit_exact_vs_oracle():
    """prob_max_ind (submodules.py:226), conf>.2 (refinenet.py:65) and x-disp>=0 (s2m2.py:179) decisions."""
    C, NTR, H, W = 64, 1, 128, 192
    sd = random_state_dict(C, NTR, 21)
    left, right = synthetic_pair(1, H, W, 22)
    m = _model(C, NTR, sd)
    m.taps = {}
    m(left.to(DEV), right.to(DEV))
    ot = {}
    _oracle_gpu(sd, left, right, NTR, True, 3, ot)
    assert torch.equal(m.taps["disp_init.argmax"], ot["disp_init.argmax"])
    assert torch.equal(m.taps["disp_init.conf"] > 0.2, ot["disp_init.conf"] > 0.2)
    for it in range(3):
        mine = m.taps[f"refiner.{it}.occ"] == 0
        ref = ot[f"refiner.{it}.occ"] == 0
        assert int((mine != ref).sum()) == 0, f"occ mask differs at iteration {it}"


@pytest.mark.parametrize("name", GOLDEN_CASES[:2] + GOLDEN_CASES[3:])
def test_fp16_path_no_worse_than_reference_autocast(name):
    """fp16 mode vs the fp32 golden, judged against the reference algorithm's OWN fp16 behaviour: the oracle under
    torch.autocast(fp16) (how run_stereo_matching runs the reference, model_utils.py:76) on the same inputs.
    With random-init weights the matching is nearly flat, so fp16 flips some arg-maxes in either implementation;
    the bound is 2x the autocast EPE (floor 5e-3 px, the BASELINE.md section-3 figure)."""
    cfg, g = load_golden(name)
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    m = _model(cfg["C"], cfg["NTR"], sd, cfg["pos"], cfg["it"], "fp16")
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    d, o, c = (t.cpu() for t in m(left.to(DEV), right.to(DEV)))
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.float16):
        ac = O.forward(sdg, left.to(DEV), right.to(DEV), cfg["NTR"], cfg["pos"], cfg["it"])
    epe = (d - g["disp"]).abs().mean().item()
    epe_ac = (ac[0].float().cpu() - g["disp"]).abs().mean().item()
    print(f"{name} fp16: EPE {epe:.3e} max {(d - g['disp']).abs().max().item():.3e} | reference-autocast EPE {epe_ac:.3e}")
    assert torch.isfinite(d).all() and torch.isfinite(o).all() and torch.isfinite(c).all()
    assert epe < max(2 * epe_ac, 5e-3)
    assert (o - g["occ"]).abs().mean() < max(2 * (ac[1].float().cpu() - g["occ"]).abs().mean().item(), 5e-3)
    assert (c - g["conf"]).abs().mean() < max(2 * (ac[2].float().cpu() - g["conf"]).abs().mean().item(), 5e-3)


@pytest.mark.parametrize("smooth", [True, False])
def test_s_640x480_vs_oracle_on_gpu(smooth):
    """BASELINE config 2 (S, 640x480): both precisions vs the oracle (fp32, TF32 off) on identical inputs — EPE < 1e-3 px and the
    integer contract bit-identical; `smooth=False` is the raw-noise input of BASELINE.md section 4."""
    C, NTR = 128, 1
    sd = random_state_dict(C, NTR, 0)
    left, right = synthetic_pair(1, 480, 640, 1, smooth=smooth)
    ot = {}
    ref = [t.cpu() for t in _oracle_gpu(sd, left, right, NTR, True, 3, ot)]
    for prec, tol in (("fp32", 1e-4), ("fp16", 1e-3)):
        m = _model(C, NTR, sd, precision=prec)
        m.taps = {}
        out = [t.cpu() for t in m(left.to(DEV), right.to(DEV))]
        flips, conf_mm, occ_mm = _decision_mismatches(m.taps, ot, 3)
        epe = (out[0] - ref[0]).abs().mean().item()
        print(f"S 640x480 {prec} smooth={smooth}: EPE {epe:.3e} max {(out[0] - ref[0]).abs().max().item():.3e} "
              f"argmax flips {flips} conf>.2 mismatches {conf_mm} occ-mask mismatches {occ_mm}")
        assert epe < tol
        assert (out[1] - ref[1]).abs().mean() < tol and (out[2] - ref[2]).abs().mean() < tol
        assert flips == 0 and conf_mm == 0 and occ_mm == [0, 0, 0]


@pytest.mark.parametrize("smooth", [True, False])
def test_l_1216x1024_headline_config_vs_oracle_on_gpu(smooth):
    """BASELINE config 3 (L, 1216x1024, refine_iter=3), the headline: both precisions vs the fp32 oracle (TF32 off): EPE < 1e-3 px,
    arg-max index / conf > .2 / per-iteration occlusion masks bit-identical — asserted for fp16 (the benchmarked mode) as well."""
    C, NTR = 256, 3
    sd = random_state_dict(C, NTR, 0)
    left, right = synthetic_pair(1, 1024, 1216, 1, smooth=smooth)
    ot = {}
    ref = [t.cpu() for t in _oracle_gpu(sd, left, right, NTR, True, 3, ot)]
    keep = ["disp_init.argmax", "disp_init.conf"] + [f"refiner.{it}.occ" for it in range(3)]
    ot = {k: ot[k].cpu() for k in keep}
    torch.cuda.empty_cache()
    for prec, tol in (("fp32", 1e-4), ("fp16", 1e-3)):
        m = _model(C, NTR, sd, precision=prec)
        m.taps = {}
        out = [t.cpu() for t in m(left.to(DEV), right.to(DEV))]
        flips, conf_mm, occ_mm = _decision_mismatches(m.taps, ot, 3)
        m.taps = None
        err = (out[0] - ref[0]).abs()
        print(f"L 1216x1024 {prec} smooth={smooth}: EPE {err.mean().item():.3e} max {err.max().item():.3e} argmax flips {flips}/"
              f"{ot['disp_init.argmax'].numel()} conf>.2 mismatches {conf_mm} occ-mask mismatches {occ_mm} "
              f"occ {(out[1] - ref[1]).abs().mean().item():.2e} conf {(out[2] - ref[2]).abs().mean().item():.2e}")
        assert torch.isfinite(out[0]).all()
        assert err.mean().item() < tol           # north_star: EPE < 1e-3 px vs the fp32 reference
        assert (out[1] - ref[1]).abs().mean() < tol and (out[2] - ref[2]).abs().mean() < tol
        assert flips == 0 and conf_mm == 0 and occ_mm == [0, 0, 0]   # north_star: bit-exact arg-max index and occlusion masks
        if prec == "fp16" and smooth:
            # BASELINE config 5's per-GPU share is a batch of pairs: pair 0 of a batch (with a different second pair) must
            # reproduce the single-pair result at the headline size (pairs are independent units)
            l2, r2 = synthetic_pair(1, 1024, 1216, 7)
            ob = [t.cpu() for t in m(torch.cat([left, l2]).to(DEV), torch.cat([right, r2]).to(DEV))]
            eb = (ob[0][0] - ref[0][0]).abs()
            db = (ob[0][0] - out[0][0]).abs()
            print(f"L 1216x1024 fp16 batch 2, pair 0: EPE vs oracle {eb.mean().item():.3e}, vs single-pair call {db.mean().item():.3e} "
                  f"max {db.max().item():.3e}")
            assert eb.mean().item() < tol and db.mean().item() < 1e-4
            assert not torch.equal(ob[0][1], ob[0][0])               # the second pair is a different problem
        del m
        torch.cuda.empty_cache()


def test_l_1216x1024_fp16_vs_real_reference_golden():
    """The headline config against the REAL reference (tests/golden/l_c256_1216x1024_raw.npz, made by oracle/make_golden_l.py
    from the unmodified module on CPU, un-softened inputs): outputs at every 4th pixel, their full-resolution means, and the
    integer contract (arg-max index, conf > .2, per-iteration occlusion masks) bit-identical in fp16 and fp32."""
    import numpy as np
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "l_c256_1216x1024_raw.npz"))
    C, NTR, B, H, W, _, IT, wseed, iseed = [int(v) for v in z["cfg"]]
    sd = random_state_dict(C, NTR, wseed)
    left, right = synthetic_pair(B, H, W, iseed, smooth=False)
    h4, w4 = H // 4, W // 4
    unpack = lambda a: torch.from_numpy(np.unpackbits(a)[:B * h4 * w4].reshape(B, 1, h4, w4).astype(bool))
    for prec, tol in (("fp32", 1e-4), ("fp16", 1e-3)):
        m = _model(C, NTR, sd, precision=prec)
        m.taps = {}
        out = [t.cpu() for t in m(left.to(DEV), right.to(DEV))]
        taps, m.taps = m.taps, None
        for name, t, mean in zip(("disp", "occ", "conf"), out, z["mean"]):
            err = (t[..., ::4, ::4] - torch.from_numpy(z[name + "_s4"])).abs()
            print(f"L golden {prec} {name}: mean err {err.mean().item():.3e} max {err.max().item():.3e} | full-res mean {t.double().mean().item():.6f} vs {mean:.6f}")
            assert err.mean().item() < tol
            assert abs(t.double().mean().item() - float(mean)) < tol
        flips = int((taps["disp_init.argmax"].cpu().reshape(B, h4, w4) != torch.from_numpy(z["argmax"].astype(np.int64)).reshape(B, h4, w4)).sum())
        conf_mm = int(((taps["disp_init.conf"].cpu().reshape(B, 1, h4, w4) > 0.2) != unpack(z["conf_mask"])).sum())
        occ_mm = [int(((taps[f"refiner.{i}.occ"].cpu().reshape(B, 1, h4, w4) != 0) != unpack(z[f"occ_mask{i}"])).sum()) for i in range(IT)]
        print(f"L golden {prec}: argmax flips {flips} conf>.2 mismatches {conf_mm} occ-mask mismatches {occ_mm}")
        assert (taps["disp_init.disp"].cpu().reshape(B, 1, h4, w4) - torch.from_numpy(z["di_disp"])).abs().max() < 2e-3
        for i in range(IT):
            assert (taps[f"refiner.{i}.disp"].cpu().reshape(B, 1, h4, w4) - torch.from_numpy(z[f"it{i}_disp"])).abs().mean() < tol
        assert flips == 0 and conf_mm == 0 and occ_mm == [0] * IT
        del m
        torch.cuda.empty_cache()


@pytest.mark.parametrize("model_type,H,W", [("M", 256, 384), ("XL", 256, 320)])
def test_other_model_sizes_vs_oracle(model_type, H, W):
    """M (CH192/NTR2) and XL (CH384/NTR3): head dims 192/96/48 and 384/192/96 — the non-power-of-two paths."""
    from s2m2_b200.weights import MODEL_TABLE
    C, NTR = MODEL_TABLE[model_type]
    sd = random_state_dict(C, NTR, 0)
    left, right = synthetic_pair(1, H, W, 1)
    ot = {}
    ref = [t.cpu() for t in _oracle_gpu(sd, left, right, NTR, True, 3, ot)]
    am_ref = ot["disp_init.argmax"].cpu()
    del ot
    # fp16 on these SMALL images (a 96 / 80-token 1/32 grid, outputs of a few px): measured 2.2e-3 px (M) with 0 arg-max flips —
    # fp16 rounding through the refiners, not a discrete flip; the north-star 1e-3 bound is asserted at the BASELINE sizes above
    for prec, tol in (("fp32", 1e-4), ("fp16", 4e-3)):
        m = _model(C, NTR, sd, precision=prec)
        m.taps = {}
        out = [t.cpu() for t in m(left.to(DEV), right.to(DEV))]
        flips = int((m.taps["disp_init.argmax"].cpu() != am_ref).sum())
        err = (out[0] - ref[0]).abs()
        print(f"{model_type} {W}x{H} {prec}: EPE {err.mean().item():.3e} max {err.max().item():.3e} argmax flips {flips}/{am_ref.numel()}")
        assert err.mean().item() < tol
        assert flips == 0
        del m
        torch.cuda.empty_cache()


def test_xl_2432x2048_config_runs_and_matches_oracle():
    """BASELINE config 4 (XL, 2432x2048: the 757 MB cost volume / N=4864 global attention case)."""
    C, NTR = 384, 3
    sd = random_state_dict(C, NTR, 0)
    left, right = synthetic_pair(1, 2048, 2432, 1)
    m = _model(C, NTR, sd, precision="fp16")
    out = [t.cpu() for t in m(left.to(DEV), right.to(DEV))]
    del m
    torch.cuda.empty_cache()
    assert all(torch.isfinite(t).all() for t in out)
    assert out[0].shape == (1, 1, 2048, 2432)
    ref = [t.cpu() for t in _oracle_gpu(sd, left, right, NTR, True, 3)]
    err = (out[0] - ref[0]).abs()
    print(f"XL 2432x2048 fp16: EPE {err.mean().item():.3e} max {err.max().item():.3e} occ {(out[1] - ref[1]).abs().mean().item():.2e}")
    assert err.mean().item() < 1e-3          # measured 4.4e-5 (round 1)
    assert (out[1] - ref[1]).abs().mean() < 1e-3 and (out[2] - ref[2]).abs().mean() < 1e-3


def test_output_upsample_branch_matches_reference_golden():
    from conftest import GOLDEN_UP2
    cfg, g = load_golden(GOLDEN_UP2)
    sd = random_state_dict(cfg["C"], cfg["NTR"], cfg["wseed"], cfg["gain"])
    left, right = synthetic_pair(cfg["B"], cfg["H"], cfg["W"], cfg["iseed"])
    for prec, tol in (("fp32", 2e-5), ("fp16", 1e-3)):       # measured 3.5e-6 / 3.3e-4
        m = S2M2(cfg["C"], 1, cfg["NTR"], use_positivity=cfg["pos"], refine_iter=cfg["it"], precision=prec, output_upsample=True).eval()
        m.my_load_state_dict(sd)
        m = m.to(DEV)
        d, o, c = (t.cpu() for t in m(left.to(DEV), right.to(DEV)))
        assert d.shape == g["disp"].shape
        print(f"output_upsample {prec}: EPE {(d - g['disp']).abs().mean().item():.3e} max {(d - g['disp']).abs().max().item():.3e}")
        assert (d - g["disp"]).abs().mean() < tol
        assert (o - g["occ"]).abs().mean() < tol and (c - g["conf"]).abs().mean() < tol


def test_batch_independence_and_determinism():
    """Pairs are independent units: a batch of two identical pairs gives two bit-identical results and repeated
    calls are bit-identical.  Against the single-pair call only fp16 rounding level agreement is required: tile shapes and
    the CUDA-core / tcgen05 choice of a layer depend on the number of output pixels, i.e. on the batch size."""
    C, NTR = 32, 1
    sd = random_state_dict(C, NTR, 0)
    left, right = synthetic_pair(1, 64, 96, 1)
    m = _model(C, NTR, sd, precision="fp16")
    l2, r2 = torch.cat([left, left]).to(DEV), torch.cat([right, right]).to(DEV)
    a = m(l2, r2)
    b = m(l2, r2)
    s = m(left.to(DEV), right.to(DEV))
    for x, y, z in zip(a, b, s):
        assert torch.equal(x, y)
        assert torch.equal(x[0], x[1])
        assert (x[0] - z[0]).abs().max() < 2e-2 and (x[0] - z[0]).abs().mean() < 1e-3, (float((x[0] - z[0]).abs().max()), float((x[0] - z[0]).abs().mean()))


def test_rejects_bad_shapes_loudly():
    m = _model(32, 1, random_state_dict(32, 1, 0))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 60, 64, device=DEV), torch.zeros(1, 3, 60, 64, device=DEV))


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.uint8])
def test_pipelined_runner_matches_direct_forward(in_dtype):
    """s2m2_b200.runtime.PipelinedRunner (CUDA graph + double-buffered H2D / D2H on a second stream): every batch of a stream of
    different inputs comes back bit-identical to the plain forward call, through run(), stream() and infer()."""
    from s2m2_b200.runtime import PipelinedRunner
    C, NTR, B, H, W = 32, 1, 2, 64, 96
    m = _model(C, NTR, random_state_dict(C, NTR, 0), precision="fp16")
    batches = []
    for k in range(5):
        l, r = synthetic_pair(B, H, W, 100 + k)
        batches.append((l.to(in_dtype).pin_memory(), r.to(in_dtype).pin_memory()))
    with torch.inference_mode():
        want = [[t.cpu() for t in m(l.to(DEV), r.to(DEV))] for l, r in batches]
    runner = PipelinedRunner(m, B, H, W, torch.device(DEV), input_dtype=in_dtype)
    outs = [[torch.empty((B, 1, H, W)).pin_memory() for _ in range(3)] for _ in batches]
    runner.run(batches, host_out=outs)
    torch.cuda.synchronize()
    for got, ref in zip(outs, want):
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    n = 0
    for got, ref in zip(runner.stream(iter(batches)), want):
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
        n += 1
    assert n == len(batches)
    for a, b in zip(runner.infer(*batches[3]), want[3]):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        runner.infer(batches[0][0][:1], batches[0][1][:1])


def test_module_auto_graph_replay_is_bit_identical_to_eager():
    """S2M2.forward replays a CUDA graph captured at the end of the first call of a (shape, dtype, crop) — the path
    run_stereo_matching's repeated calls take.  Every call must return exactly what the eager launches return, also after a shape change in between."""
    C, NTR = 32, 1
    sd = random_state_dict(C, NTR, 0)
    m = _model(C, NTR, sd, precision="fp16")
    e = _model(C, NTR, sd, precision="fp16")
    e.auto_graph = False
    pairs = [synthetic_pair(2, 64, 96, 40 + k) for k in range(4)] + [synthetic_pair(1, 96, 64, 50)] + [synthetic_pair(2, 64, 96, 60)]
    for k, (l, r) in enumerate(pairs):
        a = m(l.to(DEV), r.to(DEV))
        b = e(l.to(DEV), r.to(DEV))
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"call {k}"
    assert sum("graph" in v for v in m._graphs.values()) >= 1      # the 64x96 batch-2 shape was captured and replayed
    u8 = [(l.to(torch.uint8), r.to(torch.uint8)) for l, r in pairs[:3]]
    for l, r in u8:
        for x, y in zip(m(l.to(DEV), r.to(DEV), crop_hw=(50, 80)), e(l.to(DEV), r.to(DEV), crop_hw=(50, 80))):
            assert x.shape[-2:] == (50, 80) and torch.equal(x, y)
