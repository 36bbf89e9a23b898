"""ORACLE (test infrastructure) — restatement of the reference's pre- and post-processing around the
depth network, device-agnostic torch code that issues the same ATen ops in the same order and dtype
as the reference does.  Pinned on goldens from the unmodified reference (tests/golden/prepost.npz).

    process_cuda_branch          depth.py:542-566   (the IS_CUDA definition of process())
    model_input_shape            depth.py:676-692
    resize_patch_aligned         depth.py:676-706   (CUDA branch: bicubic + antialias)
    normalise_input              depth.py:1931, 1946-1948
    normalize / percentile       depth.py:816-867, 784-794
    apply_gamma                  depth.py:775
    apply_foreground_scale       depth.py:709-736
    anti_alias                   depth.py:740-765
    post_process_depth           depth.py:806-814
    ema                          depth.py:1865-1887
    upsample_depth               depth.py:1998-2004
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def process_cuda_branch(frame_u8: torch.Tensor, target_height: int, dtype=torch.float16) -> torch.Tensor:
    img = frame_u8[..., :3].flip(-1).permute(2, 0, 1).contiguous()
    _, H0, W0 = img.shape
    if target_height >= H0:
        return img.to(dtype)
    nh = (target_height // 2) * 2
    nw = (int(W0 * target_height / H0) // 2) * 2
    return F.interpolate(img.to(dtype).unsqueeze(0), size=(nh, nw), mode="bilinear", align_corners=False,
                         antialias=nh < H0).squeeze(0)


def model_input_shape(h: int, w: int, target: int = 518, patch: int = 14):
    longest = max(h, w)
    scale = target / float(longest) if longest != target else 1.0
    sh = max(1, int(round(h * scale)))
    sw = max(1, int(round(w * scale)))

    def nearest_multiple(x, p):
        down = (x // p) * p
        up = down + p
        return up if abs(up - x) <= abs(x - down) else down

    return max(1, nearest_multiple(sh, patch)), max(1, nearest_multiple(sw, patch))


def resize_patch_aligned(t: torch.Tensor, target: int = 518, patch: int = 14) -> torch.Tensor:
    """t [1,3,h,w] uint8 or float."""
    _, _, h, w = t.shape
    nh, nw = model_input_shape(h, w, target, patch)
    if (nh, nw) == (h, w):
        return t
    dtype = t.dtype if t.dtype.is_floating_point else torch.float32
    return F.interpolate(t.to(dtype), size=(nh, nw), mode="bicubic", align_corners=False, antialias=True)


def normalise_input(t: torch.Tensor, model_dtype=torch.float32, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    t = t.to(model_dtype) / 255.0
    m = torch.tensor(mean, dtype=model_dtype, device=t.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=model_dtype, device=t.device).view(1, 3, 1, 1)
    return (t - m) / s


def percentile_bounds(vv: torch.Tensor, percentile: float):
    vv = vv.flatten()
    n = vv.numel()
    lo_q = max(0.0, min(1.0, float(percentile) / 100.0))
    k = min(n, max(1, int(round(lo_q * (n - 1))) + 1))
    if k == n:
        return vv.min(), vv.max()
    lo = torch.topk(vv, k, largest=False, sorted=False).values
    hi = torch.topk(vv, k, largest=True, sorted=False).values
    return lo.max(), hi.min()


def normalize(depth: torch.Tensor, percentile=2.0, subsample_cap=6144, metric=False) -> torch.Tensor:
    d = depth.squeeze()
    if metric:
        valid = d > 0
        inv = torch.where(valid, 1.0 / d.clamp(min=1e-12), d)
        v = inv[valid]
    else:
        inv = d
        v = inv.flatten()
    if v.numel() <= 10:
        dmin = torch.zeros((), device=d.device)
        dmax = torch.zeros((), device=d.device)
    else:
        vv = v
        if vv.numel() > subsample_cap:
            step = (vv.numel() + subsample_cap - 1) // subsample_cap
            vv = vv[::step]
        dmin, dmax = percentile_bounds(vv, percentile)
    denom = (dmax - dmin).clamp_min(1e-6)
    return ((inv - dmin) / denom).clamp(0.0, 1.0)


def apply_gamma(depth, gamma=1.45):
    return torch.pow(depth, gamma)


def apply_foreground_scale(depth, scale: float, mid: float = 0.5, eps: float = 1e-6):
    depth = depth.clamp(0.0, 1.0)
    if abs(scale) < eps:
        return depth
    exponent = 1.0 / (1.0 + scale)
    dist = depth - mid
    out = mid + torch.sign(dist) * torch.pow(torch.abs(dist), exponent)
    return out.clamp(0.0, 1.0)


def anti_alias(depth: torch.Tensor, strength: float = 1.0) -> torch.Tensor:
    if depth.dim() == 2:
        depth = depth.unsqueeze(0).unsqueeze(0)
    k = int(3 * strength) | 1
    if k < 3:
        return depth.squeeze()
    sigma = 0.5 * strength
    coords = torch.arange(k, device=depth.device, dtype=depth.dtype) - k // 2
    gauss = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    gauss /= gauss.sum()
    depth = F.conv2d(depth, gauss.view(1, 1, 1, -1), padding=(0, k // 2))
    depth = F.conv2d(depth, gauss.view(1, 1, -1, 1), padding=(k // 2, 0))
    return depth.squeeze()


def post_process_depth(depth, foreground_scale=0.05, aa_strength=4.0, metric=False):
    depth = normalize(depth, metric=metric).squeeze()
    depth = apply_gamma(depth)
    depth = apply_foreground_scale(depth, scale=foreground_scale)
    return anti_alias(depth, strength=aa_strength)


def ema(prev, depth, alpha=0.9):
    """DepthStabilizer.__call__: returns (new_prev, output)."""
    if prev is None or prev.shape != depth.shape:
        return depth.detach().clone(), depth
    prev = prev.clone()
    prev.lerp_(depth, 1.0 - alpha)
    return prev, prev


def upsample_depth(depth, h, w):
    return F.interpolate(depth[None, None], size=(h, w), mode="bilinear", align_corners=False)[0, 0]
