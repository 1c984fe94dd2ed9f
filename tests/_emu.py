"""CPU emulation of the C-ABI operators' *addressing* (test infrastructure only).

Used by the CPU test-suite to validate the test harness, the weight packing and the padded-flat
layout arithmetic without a GPU: each function restates what the CUDA kernel does index-for-index
(flat tap offsets on the haloed buffer, im2col gather with front pads, halo zeroing), in fp32 torch
on bf16-rounded data.  Never imported by the product package.
"""
import torch

from tedspad_b200 import _lib as L


def _act(v, act):
    if act == L.ACT_RELU:
        return torch.relu(v)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(v)
    return v


def conv_forward(x, pc, y, res=None, act=L.ACT_RELU, feed=L.FEED_AUTO, y_fp32=False, max_ctas=0, n_tile=0):
    if isinstance(y, (list, tuple)):   # PackedConv.concat merge: one destination per stacked weight block
        import copy
        for yi, r0, co in zip(y, pc.seg_begin, pc.seg_cout):
            sub = copy.copy(pc)
            sub.w, sub.bias, sub.cout = pc.w[r0:r0 + co], pc.bias[r0:r0 + co], co
            conv_forward(x, sub, yi, None, act, feed, y_fp32, max_ctas, n_tile)
        return y[0]
    kd, kh, kw = pc.k
    ntaps = kd * kh * kw
    W2 = pc.w.float()[:pc.cout]  # [Cout, K_pad]
    bias = pc.bias[:pc.cout]
    flat_legal = (pc.stride == (1, 1, 1) and all(k % 2 == 1 for k in pc.k) and pc.pad_front == tuple(k // 2 for k in pc.k)
                  and x.C % 64 == 0 and (x.D, x.H, x.W) == (y.D, y.H, y.W) and x.halo == y.halo
                  and all(h >= p for h, p in zip(x.halo, pc.pad_front)) and pc.k_pad == ntaps * x.C)
    if feed == L.FEED_AUTO:
        feed = L.FEED_FLAT_TMA if flat_legal else L.FEED_GATHER
    if feed == L.FEED_FLAT_TMA:
        assert flat_legal
        Dp, Hp, Wp = x.D + 2 * x.pd, x.H + 2 * x.ph, x.W + 2 * x.pw
        M = x.N * Dp * Hp * Wp
        xf = x.buf.reshape(M, x.ld)[:, x.coff:x.coff + x.C].float()
        acc = torch.zeros(M, pc.cout)
        t = 0
        for a in range(kd):
            for b in range(kh):
                for c in range(kw):
                    off = ((a - pc.pad_front[0]) * Hp + (b - pc.pad_front[1])) * Wp + (c - pc.pad_front[2])
                    rows = torch.arange(M) + off
                    ok = (rows >= 0) & (rows < M)
                    A = torch.zeros(M, x.C)
                    A[ok] = xf[rows[ok]]
                    acc += A @ W2[:, t * x.C:(t + 1) * x.C].T
                    t += 1
        acc = acc + bias
        if res is not None:
            acc = acc + res.buf.reshape(M, res.ld)[:, res.coff:res.coff + res.C].float()
        acc = _act(acc, act)
        m = torch.arange(M)
        wq, hq, dq = m % Wp, (m // Wp) % Hp, (m // (Wp * Hp)) % Dp
        interior = (wq >= y.pw) & (wq < y.pw + y.W) & (hq >= y.ph) & (hq < y.ph + y.H) & (dq >= y.pd) & (dq < y.pd + y.D)
        acc[~interior] = 0
        y.buf.reshape(M, y.ld)[:, y.coff:y.coff + y.C] = acc.to(y.buf.dtype)
        return y
    # gather
    acc = _gather_acc(x, pc, (y.N, y.D, y.H, y.W))
    if res is not None:
        acc = acc + res.interior().float()
    acc = _act(acc, act)
    y.interior()[...] = acc.to(y.buf.dtype)
    return y


def _gather_acc(x, pc, out_shape):
    """fp32 conv + bias by explicit im2col with front pads: [N,OD,OH,OW,Cout] (x may carry fewer channels than
    the weights were packed for, e.g. the 4-channel clip of the SLAB stem: missing channels are zero)."""
    kd, kh, kw = pc.k
    W2 = pc.w.float()[:pc.cout]
    bias = pc.bias[:pc.cout]
    N, OD, OH, OW = out_shape
    xi = x.interior().float()  # [N,D,H,W,C]
    Cw = pc.cin_pad
    if x.C < Cw:
        xi = torch.cat([xi, torch.zeros(*xi.shape[:-1], Cw - x.C)], -1)
    cols = torch.zeros(N, OD, OH, OW, pc.k_pad)
    t = 0
    od = torch.arange(OD) * pc.stride[0] - pc.pad_front[0]
    oh = torch.arange(OH) * pc.stride[1] - pc.pad_front[1]
    ow = torch.arange(OW) * pc.stride[2] - pc.pad_front[2]
    for a in range(kd):
        for b in range(kh):
            for c in range(kw):
                idd, ihh, iww = od + a, oh + b, ow + c
                vd, vh, vw = (idd >= 0) & (idd < x.D), (ihh >= 0) & (ihh < x.H), (iww >= 0) & (iww < x.W)
                patch = torch.zeros(N, OD, OH, OW, Cw)
                sub = xi[:, idd[vd]][:, :, ihh[vh]][:, :, :, iww[vw]]
                tmp = torch.zeros(N, int(vd.sum()), int(vh.sum()), OW, Cw)
                tmp[:, :, :, vw] = sub
                tmp2 = torch.zeros(N, int(vd.sum()), OH, OW, Cw)
                tmp2[:, :, vh] = tmp
                patch[:, vd] = tmp2
                cols[..., t * Cw:(t + 1) * Cw] = patch
                t += 1
    acc = cols.reshape(-1, pc.k_pad) @ W2.T + bias
    return acc.reshape(N, OD, OH, OW, pc.cout)


def slab_runs_kx(x, psc):
    return psc is not None and psc.kind == L.SLAB_3X3_KX_PAIR and psc.resolve(x) is psc


def conv_slab_forward(x, psc, y, act=L.ACT_RELU, pool=None, outconv=None, tm=0, max_ctas=0, up=None, stack_rows=0, res=None,
                      s2d_clip=None):
    """SLAB feed (csrc/conv_slab.cu): same arithmetic as the gather restatement; the fused MaxPool2d(2) pools the
    ROUNDED output (as the kernel does), the fused OutConv consumes the un-rounded fp32 activations; with `up` the
    convolution input is [x | upsample2x(up)] (the up-sampled half rounded to the storage dtype, as the kernel does)."""
    pc = psc.pc
    if up is not None:
        from tedspad_b200.ops import CLTensor
        cat = CLTensor(x.N, x.D, x.H, x.W, x.C + up.C, device="cpu", dtype=x.buf.dtype)
        cat.interior()[..., :x.C] = x.interior()
        upsample2x(up, cat.slice(x.C, up.C))
        x = cat
    shape = (y.N, y.D, y.H, y.W) if y is not None else (x.N, x.D, x.H, x.W)
    acc = _gather_acc(x, pc, shape)
    if res is not None:
        acc = acc + res.interior().float()
    acc = _act(acc, act)
    store_dtype = y.buf.dtype if y is not None else (pool.buf.dtype if pool is not None else torch.float32)
    if s2d_clip is not None:   # KX kind: 12 space-to-depth channels -> encoder clip (rounded to the storage dtype first)
        clip, T = s2d_clip
        v = acc.to(clip.buf.dtype).float()[:, 0, :, :, :12]
        F_, hh, ww, _ = v.shape
        fr = v.reshape(F_, hh, ww, 2, 2, 3).permute(0, 5, 1, 3, 2, 4).reshape(F_, 3, 2 * hh, 2 * ww).contiguous()
        planes_to_clip_into(fr, clip, T)
    if y is not None:
        y.interior()[...] = acc.to(y.buf.dtype)
    if pool is not None:
        import torch.nn.functional as F
        r = acc.to(store_dtype).float()[:, 0].permute(0, 3, 1, 2)
        pool.interior()[:, 0] = F.max_pool2d(r, 2).permute(0, 2, 3, 1).to(pool.buf.dtype)
    if outconv is not None:
        w, b, planes, frames = outconv[:4]
        v = torch.sigmoid(acc[:, 0] @ w.T + b).permute(0, 3, 1, 2).contiguous()  # [N,3,H,W]
        if planes is not None:
            planes.copy_(v.to(planes.dtype))
        if frames is not None:
            frames.copy_(v)
        if len(outconv) > 4:   # straight into the encoder clip through the raw-reshape glue
            clip, T = outconv[4], outconv[5]
            planes_to_clip_into(v, clip, T)
    return y


def planes_to_clip_into(v, y, T):
    """fp32 [F,3,H,W] images -> channels 0..2 of the encoder clip (other channels untouched)."""
    Fr, _, H, W = v.shape
    B = Fr // T
    enc = v.float().reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)  # dali_extraction.py:171-173
    y.interior()[..., :3] = enc.permute(0, 2, 3, 4, 1).to(y.buf.dtype)
    return y


def planes_to_clip(planes, y, T):
    Fr, _, H, W = planes.shape
    B = Fr // T
    enc = planes.float().reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)  # dali_extraction.py:171-173
    y.interior()[...] = 0
    y.interior()[..., :3] = enc.permute(0, 2, 3, 4, 1).to(y.buf.dtype)
    return y


def maxpool(x, y, k, s, pad_front=(0, 0, 0), zero_pad=False):
    import torch.nn.functional as F
    xi = x.interior().float().permute(0, 4, 1, 2, 3)
    need = [(o - 1) * st + kk - pf - i for o, st, kk, pf, i in zip((y.D, y.H, y.W), s, k, pad_front, (x.D, x.H, x.W))]
    pb = [max(n, 0) for n in need]
    xp = F.pad(xi, (pad_front[2], pb[2], pad_front[1], pb[1], pad_front[0], pb[0]),
               value=0.0 if zero_pad else float("-inf"))
    out = F.max_pool3d(xp, k, s)[:, :, :y.D, :y.H, :y.W]
    y.interior()[...] = out.permute(0, 2, 3, 4, 1).to(y.buf.dtype)
    return y


def upsample2x(x, y):
    import torch.nn.functional as F
    xi = x.interior().float()[:, 0].permute(0, 3, 1, 2)
    up = F.interpolate(xi, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = y.H - up.shape[2], y.W - up.shape[3]
    up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    y.interior()[:, 0] = up.permute(0, 2, 3, 1).to(y.buf.dtype)
    return y


def upsample2x_nearest(x, y):
    import torch.nn.functional as F
    xi = x.interior().float()[:, 0].permute(0, 3, 1, 2)
    up = F.interpolate(xi, scale_factor=2, mode="nearest")
    y.interior()[:, 0] = up.permute(0, 2, 3, 1).to(y.buf.dtype)
    return y


def frames_to_clip(x, y, T, frames_out=None, s2d=False):
    if s2d:   # [F,h,w,(a,b,c)] -> [F,3,2h,2w]
        v = x.interior().float()[:, 0, :, :, :12]
        F_, h, w, _ = v.shape
        fr = v.reshape(F_, h, w, 2, 2, 3).permute(0, 5, 1, 3, 2, 4).reshape(F_, 3, 2 * h, 2 * w).contiguous()
    else:
        fr = x.interior().float()[:, 0, :, :, :3].permute(0, 3, 1, 2).contiguous()   # [F,3,H,W]
    if frames_out is not None:
        frames_out.copy_(fr)
    y.interior()[...] = 0
    return planes_to_clip_into(fr, y, T)


def outconv_sigmoid(x, w, b, y, T, frames_out=None):
    xi = x.interior().float()[:, 0]  # [F,H,W,C]
    v = torch.sigmoid(xi @ w.T + b)  # [F,H,W,3]
    Fr, H, W, _ = v.shape
    B = Fr // T
    fr = v.permute(0, 3, 1, 2).contiguous()  # [F,3,H,W]
    if frames_out is not None:
        frames_out.copy_(fr)
    enc = fr.reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)
    y.interior()[..., :3] = enc.permute(0, 2, 3, 4, 1).to(y.buf.dtype)
    return y


def avgpool_features(x, kd=0):
    xi = x.interior().float()
    kd = kd if kd > 0 else x.D
    outs = [xi[:, d:d + kd].mean(dim=(1, 2, 3)) for d in range(x.D - kd + 1)]
    return torch.stack(outs, 1)


def nchw_to_cl(x_f32, y):
    if x_f32.dim() == 4:
        x_f32 = x_f32.unsqueeze(2)
    assert y.C % 8 == 0 or y.C == 4
    y.interior()[...] = 0
    y.interior()[..., :x_f32.shape[1]] = x_f32.permute(0, 2, 3, 4, 1).to(y.buf.dtype)
    return y


def preprocess(frames_u8, desc_i32, crop_hw, y, resample=L.RESAMPLE_AA_FLOAT, frames_f32=None):
    import numpy as np
    from oracle import preprocess as P
    fr = frames_u8.cpu().numpy()
    ch, cw = crop_hw
    outs = []
    for s, t, l, fl in desc_i32.cpu().tolist():
        if s < 0:
            outs.append(np.zeros((3, y.H, y.W), np.float32))
            continue
        img = fr[s][:, ::-1] if fl else fr[s]
        crop = img[t:t + ch, l:l + cw]
        if resample == L.RESAMPLE_PIL_U8:
            o = P.resize_pil_u8(crop, y.H, y.W).astype(np.float32).transpose(2, 0, 1) / np.float32(255)
        else:
            o = P.resize_aa(crop.astype(np.float32).transpose(2, 0, 1) / np.float32(255), y.H, y.W)
        outs.append(o)
    o = torch.from_numpy(np.stack(outs))
    if frames_f32 is not None:
        frames_f32.copy_(o)
    y.interior()[...] = 0
    y.interior()[:, 0, :, :, :3] = o.permute(0, 2, 3, 1).to(y.buf.dtype)
    return y
