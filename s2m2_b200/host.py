"""ctypes binding of the handle-level C-ABI (``s2m2b200_create / set_param / finalize_params / workspace_bytes / forward /
destroy``, include/s2m2b200.h) — the forward with its host side in C++ (``csrc/host_forward.cu``).

This is what a non-Python host binds (INTEGRATION.md shows the same calls from C); from Python it is the second way to run the
network, next to ``model.S2M2`` (the nn.Module surface the reference's own callers use).  Both run the same kernels with the
same packed weights; ``tests/test_host_forward_gpu.py`` checks that their outputs are bit-identical.

torch appears here only as the owner of the device buffers handed to the C side as raw pointers.
"""
from __future__ import annotations

import ctypes as C

import torch

from .ops import F16, F32, load_library


class HostForward:
    """``S2M2(feature_channels, 1, num_transformer, ...)`` behind the C handle.  ``load_state_dict`` takes the reference's
    state dict (s2m2.py:69-78 keys); calling the object runs the forward on the current CUDA stream."""

    def __init__(self, feature_channels: int, num_transformer: int = 1, use_positivity: bool = False, output_upsample: bool = False,
                 refine_iter: int = 3, precision: str = "fp16", graph: bool = False):
        self.lib = load_library()
        lib = self.lib
        lib.s2m2b200_workspace_bytes.restype = C.c_int64
        lib.s2m2b200_stereo_match_workspace_bytes.restype = C.c_int64
        lib.s2m2b200_packed_count.restype = C.c_int64
        lib.s2m2b200_destroy.restype = None
        self.output_upsample = bool(output_upsample)
        self._h = C.c_void_p()
        rc = lib.s2m2b200_create(int(feature_channels), int(num_transformer), int(bool(use_positivity)), int(self.output_upsample), int(refine_iter),
                                 F16 if precision == "fp16" else F32, C.byref(self._h))
        self._chk(rc, "create")
        if graph:      # the C side captures each shape's launches once per set of buffer addresses and replays
            self._chk(lib.s2m2b200_set_graph_mode(self._h, 1), "set_graph_mode")
        self.graph = bool(graph)
        self._io = {}
        self._work = None
        self._planned = {}

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"s2m2b200_{what} failed ({rc}): {self.lib.s2m2b200_last_error().decode()}")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.s2m2b200_destroy(h)

    # ---- parameters ---------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        for name, t in state_dict.items():
            t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._chk(self.lib.s2m2b200_set_param(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()), f"set_param({name})")
        self._chk(self.lib.s2m2b200_finalize_params(self._h), "finalize_params")
        return self

    def packed_entries(self):
        """[(name, ctypes address, bytes, dtype, on_device)] of the packed blob (tests)."""
        out = []
        for i in range(int(self.lib.s2m2b200_packed_count(self._h))):
            name, ptr, nbytes, dt, dev = C.c_char_p(), C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
            self._chk(self.lib.s2m2b200_packed_entry(self._h, C.c_int64(i), C.byref(name), C.byref(ptr), C.byref(nbytes), C.byref(dt), C.byref(dev)),
                      "packed_entry")
            out.append((name.value.decode(), ptr.value, nbytes.value, dt.value, bool(dev.value)))
        return out

    # ---- forward ------------------------------------------------------------------------------------
    def workspace_bytes(self, B, H, W, in_u8=False, crop_hw=None) -> int:
        ch, cw = (int(crop_hw[0]), int(crop_hw[1])) if crop_hw is not None else (0, 0)
        n = int(self.lib.s2m2b200_workspace_bytes(self._h, B, H, W, int(in_u8), ch, cw))
        if n < 0:
            self._chk(-1, "workspace_bytes")
        return n

    @torch.no_grad()
    def __call__(self, img0: torch.Tensor, img1: torch.Tensor, crop_hw=None):
        """(B,3,H,W) fp32 [0,255] or uint8 CUDA images -> disparity, occlusion, confidence, each (B,1,Ho,Wo) fp32."""
        if img0.shape != img1.shape or img0.dim() != 4 or img0.shape[1] != 3 or not img0.is_cuda:
            raise ValueError("expected two CUDA (B,3,H,W) images of equal shape")
        u8 = img0.dtype == torch.uint8 and img1.dtype == torch.uint8
        in0, in1 = (img0.contiguous(), img1.contiguous()) if u8 else (img0.float().contiguous(), img1.float().contiguous())
        B, _, H, W = in0.shape
        ch, cw = (int(crop_hw[0]), int(crop_hw[1])) if crop_hw is not None and not self.output_upsample else (0, 0)
        key = (B, H, W, u8, ch, cw, str(in0.device))
        with torch.cuda.device(in0.device):
            if key not in self._planned:
                self._planned[key] = self.workspace_bytes(B, H, W, u8, (ch, cw) if ch else None)
            need = self._planned[key]
            if self._work is None or self._work.numel() < need or self._work.device != in0.device:
                self._work = torch.empty(need, dtype=torch.uint8, device=in0.device)
            sc = 2 if self.output_upsample else 1
            ho, wo = (ch, cw) if ch else (sc * H, sc * W)
            if self.graph:     # stable addresses for the captured graph: private input / output buffers per shape
                if key not in self._io:
                    self._io[key] = (torch.empty_like(in0), torch.empty_like(in1),
                                     torch.empty((3, B, 1, ho, wo), dtype=torch.float32, device=in0.device))
                g0, g1, out = self._io[key]
                g0.copy_(in0); g1.copy_(in1)
                in0, in1 = g0, g1
            else:
                out = torch.empty((3, B, 1, ho, wo), dtype=torch.float32, device=in0.device)
            rc = self.lib.s2m2b200_forward(self._h, C.c_void_p(in0.data_ptr()), C.c_void_p(in1.data_ptr()), int(u8), B, H, W, ch, cw,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(self._work.data_ptr()), C.c_int64(self._work.numel()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._chk(rc, "forward")
        if self.graph:
            out = out.clone()          # the graph's output buffer is overwritten by the next call
        return out[0], out[1], out[2]

    @torch.no_grad()
    def stereo_match(self, left: torch.Tensor, right: torch.Tensor, margin: int = 100):
        """``run_stereo_matching`` (model_utils.py:50-95) through ``s2m2b200_stereo_match``: (B,3,H0,W0) uint8 / fp32 CUDA images of
        any size >= 32 -> disparity, occlusion, confidence as (B,H0,W0) maps and the (B,) confidence scores, all on the device."""
        if left.shape != right.shape or left.dim() != 4 or left.shape[1] != 3 or not left.is_cuda:
            raise ValueError("expected two CUDA (B,3,H,W) images of equal shape")
        u8 = left.dtype == torch.uint8 and right.dtype == torch.uint8
        in0, in1 = (left.contiguous(), right.contiguous()) if u8 else (left.float().contiguous(), right.float().contiguous())
        B, _, H0, W0 = in0.shape
        with torch.cuda.device(in0.device):
            need = int(self.lib.s2m2b200_stereo_match_workspace_bytes(self._h, B, H0, W0))
            if need < 0:
                self._chk(-1, "stereo_match_workspace_bytes")
            if self._work is None or self._work.numel() < need or self._work.device != in0.device:
                self._work = torch.empty(need, dtype=torch.uint8, device=in0.device)
            out = torch.empty((3, B, H0, W0), dtype=torch.float32, device=in0.device)
            score = torch.empty((B,), dtype=torch.float32, device=in0.device)
            rc = self.lib.s2m2b200_stereo_match(self._h, C.c_void_p(in0.data_ptr()), C.c_void_p(in1.data_ptr()), int(u8), B, H0, W0, int(margin),
                                                C.c_void_p(out.data_ptr()), C.c_void_p(score.data_ptr()), C.c_void_p(self._work.data_ptr()),
                                                C.c_int64(self._work.numel()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._chk(rc, "stereo_match")
        return out[0], out[1], out[2], score
