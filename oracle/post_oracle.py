"""TEST INFRASTRUCTURE — CPU restatement of the reference's post-processing around the forward.

Only tests/ may import this.  Each function cites the reference lines it follows.
"""
import torch


def image_pad(img: torch.Tensor, factor: int = 32) -> torch.Tensor:
    """image_utils.py:27-63: centre zero-padding to multiples of ``factor`` (floor(pad/2) in front), the frame then replaced by
    adaptive_avg_pool2d(padded, [H // factor, W // factor]) -> bilinear up to the padded size, and the original pasted back."""
    import math

    import torch.nn.functional as F
    H, W = img.shape[-2:]
    Hn, Wn = math.ceil(H / factor) * factor, math.ceil(W / factor) * factor
    ph, pw = Hn - H, Wn - W
    x = F.pad(img.float(), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), "constant", 0)
    out = F.interpolate(F.adaptive_avg_pool2d(x, output_size=[H // factor, W // factor]), size=[Hn, Wn], mode="bilinear")
    out[:, :, ph // 2: ph // 2 + H, pw // 2: pw // 2 + W] = img.float()
    return out


def remap_u8(src, mapx, mapy):
    """cv2.remap(src uint8 (H,W,3), mapx, mapy float32, INTER_LINEAR, BORDER_CONSTANT 0) restated in numpy — what
    image_utils.py:108-136 (rectify_images) computes.  OpenCV's 8-bit bilinear remap is fixed point (imgproc remap.cpp /
    imgwarp initInterTab2D): coordinates rounded half-even to 1/32 pixel, weights (32-fy)(32-fx)*32 ... on a 2^15 scale with the
    integer-coordinate entry [32767, 0, 0, 1], result (sum + 2^14) >> 15, out-of-image taps read 0.  tests/test_post_cpu.py pins this
    against cv2.remap itself (cv2 is in the image) and tests/golden/remap_u8.npz holds cv2's output for the GPU test."""
    import numpy as np
    H, W, _ = src.shape
    sx = np.rint(mapx.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31
    w = np.stack([(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32], -1)
    w[(fx == 0) & (fy == 0)] = np.array([32767, 0, 0, 1])
    acc = np.zeros(mapx.shape + (3,), np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        y, x = iy + dy, ix + dx
        ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
        v = np.where(ok[..., None], src[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)].astype(np.int64), 0)
        acc += v * w[..., k][..., None]
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def image_crop(img: torch.Tensor, img_shape):
    """image_utils.py:66-96: centre crop of the last two dims, floor(d/2) dropped in front, the rest behind."""
    H, W = img.shape[-2:]
    Hn, Wn = img_shape
    ch, cw = H - Hn, W - Wn
    if ch > 0:
        img = img[..., ch // 2: H - (ch - ch // 2), :]
    if cw > 0:
        img = img[..., :, cw // 2: W - (cw - cw // 2)]
    return img


def confidence_score(pred_conf: torch.Tensor, img_hw, margin: int = 100) -> torch.Tensor:
    """model_utils.py:86-95: crop to the un-padded size, squeeze, ``[margin:-margin, margin:-margin].mean()`` — per pair."""
    c = image_crop(pred_conf.float(), img_hw)
    if c.dim() == 4:
        c = c[:, 0]
    return c[:, margin:-margin, margin:-margin].double().mean(dim=(1, 2)).float()


def disparity_to_points(disp, fx, cx, cy, baseline, doffs=0.0, conf=None, occ=None, depth_scale=1000.0, depth_trunc=1e9):
    """model_utils.py:111-136 (depth = baseline*fx/(disp+doffs), depth[disp<=0] = 1e9) followed by what open3d's
    ``RGBDImage.create_from_color_and_depth`` / ``create_from_rgbd_image`` do with it: z = depth/depth_scale, depths beyond
    ``depth_trunc`` dropped, x = (u-cx) z/fx, y = (v-cy) z/fy (fy = fx here).  Trust mask: vis_utils.py:62."""
    d = disp.float()
    if d.dim() == 4:
        d = d[:, 0]
    B, H, W = d.shape
    depth = baseline * fx / (d + doffs)
    depth = torch.where(d <= 0, torch.full_like(depth, 1e9), depth)
    z = depth / depth_scale
    ok = (z > 0) & (z <= depth_trunc)
    z = torch.where(ok, z, torch.zeros_like(z))
    u = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    v = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    xyz = torch.stack([(u - cx) * z / fx, (v - cy) * z / fx, z], dim=-1)
    valid = ok
    if conf is not None:
        valid = valid & (conf.float().reshape(B, H, W) > 0.1)
    if occ is not None:
        valid = valid & (occ.float().reshape(B, H, W) > 0.5)
    return xyz, valid
