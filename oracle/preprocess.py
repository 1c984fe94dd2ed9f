"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement of the two input stages of the reference extraction path:

  * `dali_val_augmentations`  - DALIDataloader.val_augmentations,
    /root/reference/feature_extraction/dali_extraction.py:38-50: FHWC float 0..255 -> CHW, /255,
    torchvision center_crop (offsets int(round((H-ch)/2)), torchvision functional.py:592-593), then
    F.resize(antialias=True) = aten::_upsample_bilinear2d_aa, align_corners=False.
  * `shanghai_augmentation`   - shanghai_frames_dataset.augmentation,
    /root/reference/feature_extraction/shanghai_dl.py:27-40: to_pil_image -> center_crop(int(H*0.8),
    int(H*0.8)) -> PIL BILINEAR resize (8-bit two-pass, re-quantised) -> to_tensor (/255).

Both are written out explicitly in numpy (no torchvision/PIL calls) so that the CUDA kernel can be
checked tap-for-tap; tests/test_oracle.py pins them against torchvision and Pillow themselves.
Third-party arithmetic restated here: Pillow `ImagingResample` 8bpc path (libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc/Vertical_8bpc;
PRECISION_BITS = 32-8-2), reference pin Pillow 9.4.0, container Pillow 12.2.0 (same algorithm).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def center_crop_offsets(h, w, ch, cw):
    """torchvision.transforms.functional.center_crop offsets (functional.py:592-593)."""
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def crop_size(h, w, cropping_factor=0.8, no_ar_distortion=False, square_from_h=False):
    """dali_extraction.py:45-48 (and shanghai_dl.py:32-35, which uses H for both sides)."""
    if no_ar_distortion:
        m = min(h, w)
        return int(m * cropping_factor), int(m * cropping_factor)
    if square_from_h:
        return int(h * cropping_factor), int(h * cropping_factor)
    return int(h * cropping_factor), int(w * cropping_factor)


# ----------------------------------------------------------------------------- aten antialias
def aa_axis_weights(in_size, out_size, dtype=np.float32):
    """Per-output (lo, weights) of aten's antialiased bilinear filter (UpSampleKernel.cpp
    HelperInterpLinear / upsample_bilinear2d_aa CUDA kernel), computed in `dtype`."""
    f = dtype
    scale = f(in_size) / f(out_size)
    support = scale if scale >= 1 else f(1.0)
    invscale = f(1.0) / scale if scale >= 1 else f(1.0)
    out = []
    for i in range(out_size):
        center = scale * (f(i) + f(0.5))
        lo = max(int(center - support + f(0.5)), 0)
        hi = min(int(center + support + f(0.5)), in_size)
        w = np.zeros(hi - lo, dtype=f)
        for j in range(hi - lo):
            a = abs((f(j + lo) - center + f(0.5)) * invscale)
            w[j] = f(1.0) - a if a < 1 else f(0.0)
        tot = w.sum(dtype=f)
        if tot != 0:
            w = w / tot
        out.append((lo, w))
    return out


def resize_aa(img, out_h, out_w, dtype=np.float32):
    """img: float [..., H, W] -> [..., out_h, out_w]; separable antialiased bilinear."""
    img = np.asarray(img, dtype=dtype)
    H, W = img.shape[-2:]
    wx = aa_axis_weights(W, out_w, dtype)
    wy = aa_axis_weights(H, out_h, dtype)
    tmp = np.zeros(img.shape[:-1] + (out_w,), dtype=dtype)
    for i, (lo, w) in enumerate(wx):
        tmp[..., i] = (img[..., lo:lo + len(w)] * w).sum(-1, dtype=dtype)
    out = np.zeros(img.shape[:-2] + (out_h, out_w), dtype=dtype)
    for i, (lo, w) in enumerate(wy):
        out[..., i, :] = (tmp[..., lo:lo + len(w), :] * w[:, None]).sum(-2, dtype=dtype)
    return out


def dali_val_augmentations(video_fhwc, reso=(224, 224), cropping_factor=0.8, no_ar_distortion=False):
    """video_fhwc: [T,H,W,3] values 0..255 (uint8 or float) -> float32 [T,3,reso_h,reso_w] in [0,1]."""
    v = np.asarray(video_fhwc, dtype=np.float32).transpose(0, 3, 1, 2) / np.float32(255.0)
    H, W = v.shape[-2:]
    ch, cw = crop_size(H, W, cropping_factor, no_ar_distortion)
    top, left = center_crop_offsets(H, W, ch, cw)
    v = v[..., top:top + ch, left:left + cw]
    return resize_aa(v, reso[0], reso[1])


def dali_crop_augmentations(video_fhwc, box, crop_hw, reso=(224, 224)):
    """val_augmentations with an explicit crop instead of the centre one: box = (top, left, hflip) in torchvision
    five_crop / ten_crop terms (the crop is taken from the horizontally flipped frame when hflip), crop_hw = the
    reference's crop size (dali_extraction.py:45-48).  box = the centre box reproduces dali_val_augmentations."""
    v = np.asarray(video_fhwc, dtype=np.float32).transpose(0, 3, 1, 2) / np.float32(255.0)
    top, left, hflip = box
    if hflip:
        v = v[..., ::-1]
    v = v[..., top:top + crop_hw[0], left:left + crop_hw[1]]
    return resize_aa(np.ascontiguousarray(v), reso[0], reso[1])


# ----------------------------------------------------------------------------- Pillow 8-bit
def pil_axis_coeffs(in_size, out_size):
    """Pillow precompute_coeffs (bilinear, support 1.0) + normalize_coeffs_8bpc -> [(xmin, int kk[])]."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = np.zeros(xmax, dtype=np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            k[x] = 1.0 - a if a < 1.0 else 0.0
        ww = k.sum()
        if ww != 0.0:
            k = k / ww
        kk = np.array([int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
                       for v in k], dtype=np.int64)
        out.append((xmin, kk))
    return out


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255)


def resize_pil_u8(img_hwc, out_h, out_w):
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C]: horizontal pass then vertical pass, each rounded to 8 bit."""
    img = np.asarray(img_hwc).astype(np.int64)
    H, W, C = img.shape
    if W != out_w:
        tmp = np.zeros((H, out_w, C), dtype=np.int64)
        for xx, (xmin, kk) in enumerate(pil_axis_coeffs(W, out_w)):
            acc = (1 << (PRECISION_BITS - 1)) + (img[:, xmin:xmin + len(kk), :] * kk[None, :, None]).sum(1)
            tmp[:, xx, :] = _clip8(acc)
        img = tmp
    if H != out_h:
        out = np.zeros((out_h, img.shape[1], C), dtype=np.int64)
        for yy, (ymin, kk) in enumerate(pil_axis_coeffs(H, out_h)):
            acc = (1 << (PRECISION_BITS - 1)) + (img[ymin:ymin + len(kk), :, :] * kk[:, None, None]).sum(0)
            out[yy] = _clip8(acc)
        img = out
    return img.astype(np.uint8)


def shanghai_augmentation(frame_hwc_u8, reso=(224, 224), cropping_factor=0.8, no_ar_distortion=False):
    """One decoded (BGR) frame uint8 [H,W,3] -> float32 [3,reso_h,reso_w] in [0,1] (shanghai_dl.py:27-40)."""
    H, W = frame_hwc_u8.shape[:2]
    # shanghai_dl.py:30 takes min(image.shape) over (H, W, 3) = 3 when no_ar_distortion; default path uses H twice
    if no_ar_distortion:
        m = min(frame_hwc_u8.shape)
        ch = cw = int(m * cropping_factor)
    else:
        ch, cw = crop_size(H, W, cropping_factor, square_from_h=True)
    top, left = center_crop_offsets(H, W, ch, cw)
    crop = frame_hwc_u8[top:top + ch, left:left + cw]
    out = resize_pil_u8(crop, reso[0], reso[1])
    return out.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)


# ----------------------------------------------------------------------------- multi-crop
def multi_crop_boxes(h, w, ch, cw, ncrops):
    """(top, left, hflip) per crop in torchvision order: five_crop = tl, tr, bl, br, center
    (functional.py:812-819), ten_crop = those five then the same five of the h-flipped image
    (functional.py:857-865).  ncrops == 1 is the reference's single centre crop."""
    ct, cl = center_crop_offsets(h, w, ch, cw)
    if ncrops == 1:
        return [(ct, cl, 0)]
    five = [(0, 0), (0, w - cw), (h - ch, 0), (h - ch, w - cw), (ct, cl)]
    boxes = [(t, l, 0) for t, l in five]
    if ncrops == 5:
        return boxes
    if ncrops == 10:
        return boxes + [(t, l, 1) for t, l in five]
    raise ValueError("ncrops must be 1, 5 or 10")
