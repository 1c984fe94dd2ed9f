"""GPU diagnostic battery (not a pytest file): runs every C-ABI operator against the matching
torch fp32 op on bf16-rounded operands and prints detailed error statistics, continuing past
failures.  Usage on the GPU box:

    python tests/gpu_diag.py <group> [...]      groups: flat gather ops prep perf

Each group should be run in its own process (a trapped kernel poisons the CUDA context).
"""
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tedspad_b200 import ops, _lib as L  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
RESULTS = []


def bf(t):
    return t.to(torch.bfloat16).float()


def report(name, got, ref, tol_rel=2e-2, extra=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    mx = err.max().item()
    ok = bool(torch.isfinite(got).all()) and mx <= tol_rel * scale
    RESULTS.append((name, ok))
    print(f"[{'PASS' if ok else 'FAIL'}] {name}: max_abs_err={mx:.4e} ref_max={scale:.4e} rel={mx / scale:.3e} "
          f"mean_err={err.mean().item():.3e} {extra}", flush=True)
    if not ok:
        bad = err > tol_rel * scale
        print(f"       bad elements: {int(bad.sum())}/{bad.numel()}  nan={int(torch.isnan(got).sum())}")
        idx = bad.nonzero()[:8]
        for i in idx:
            i = tuple(int(v) for v in i)
            print(f"       at {i}: got={got[i].item():.5f} ref={ref[i].item():.5f}")
        if got.dim() == 5:  # N C D H W: error by channel mod 16, by w mod 8
            e = err.amax(dim=(0, 2, 3))  # C, W
            print("       max err by channel%16:", [f"{e[c::16].max().item():.2e}" for c in range(min(16, e.shape[0]))])
            ew = err.amax(dim=(0, 1, 2, 3))
            print("       max err by w (first 24):", [f"{v:.1e}" for v in ew[:24].tolist()])
            eh = err.amax(dim=(0, 1, 2, 4))
            print("       max err by h (first 24):", [f"{v:.1e}" for v in eh[:24].tolist()])
    return ok


def conv_ref(x, w, b, stride, pad6, res=None, act="relu"):
    """x [N,C,D,H,W] fp32 (already bf16-rounded), w [Cout,Cin,kd,kh,kw]; pad6 = F.pad order (wl,wr,hl,hr,dl,dr)."""
    y = F.conv3d(F.pad(x, pad6), w, b, stride=stride)
    if res is not None:
        y = y + res
    if act == "relu":
        y = torch.relu(y)
    elif act == "sigmoid":
        y = torch.sigmoid(y)
    return y


def run_conv_case(name, N, dhw, cin, cout, k, stride=(1, 1, 1), pad_f=None, pad_b=None, halo=(0, 0, 0), feed=L.FEED_AUTO,
                  use_res=False, act="relu", cin_real=None, in_ld=None, in_coff=0, out_ld=None, out_coff=0,
                  y_fp32=False, max_ctas=0, seed=0, bn=True):
    try:
        g = torch.Generator(device="cpu").manual_seed(seed)
        D, H, W = dhw
        cin_real = cin_real or cin
        pad_f = pad_f if pad_f is not None else tuple(kk // 2 for kk in k)
        pad_b = pad_b if pad_b is not None else pad_f
        x = torch.randn(N, cin_real, D, H, W, generator=g).to(DEV)
        w = (torch.randn(cout, cin_real, *k, generator=g) / (cin_real * k[0] * k[1] * k[2]) ** 0.5).to(DEV)
        b = (torch.rand(cout, generator=g) - 0.5).to(DEV)
        bnp = None
        if bn:
            bnp = ((torch.rand(cout, generator=g) + 0.5).to(DEV), (torch.rand(cout, generator=g) - 0.5).to(DEV),
                   (torch.rand(cout, generator=g) - 0.5).to(DEV), (torch.rand(cout, generator=g) + 0.5).to(DEV), 1e-3)
        pc = ops.PackedConv(w, b, bnp, stride=stride, pad_front=pad_f, cin_pad=cin, device=DEV)
        od, oh, ow = pc.out_extent((D, H, W), pad_b)
        # input buffer (possibly a channel slice of a wider buffer, filled with junk elsewhere)
        ld_in = in_ld or cin
        xb = ops.CLTensor(N, D, H, W, ld_in, halo, device=DEV)
        if ld_in != cin:
            xb.interior()[...] = 7.0  # junk in the channels outside the view must not leak
        xv = xb.slice(in_coff, cin)
        xv.interior()[..., :cin_real] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
        if cin_real < cin:
            xv.interior()[..., cin_real:] = 0
        out_halo = halo if (od, oh, ow) == (D, H, W) else (0, 0, 0)
        ld_out = out_ld or cout
        yb = ops.CLTensor(N, od, oh, ow, ld_out, out_halo, device=DEV, dtype=torch.float32 if y_fp32 else torch.bfloat16)
        yb.buf.fill_(3.0)  # poison: halo must come back as zeros in FLAT feed, untouched in GATHER feed
        yv = yb.slice(out_coff, cout)
        resv, res_t = None, None
        if use_res:
            res_t = torch.randn(N, cout, od, oh, ow, generator=g).to(DEV)
            rb = ops.CLTensor(N, od, oh, ow, cout, out_halo, device=DEV)
            rb.interior()[...] = res_t.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
            resv = rb
        acts = {"relu": L.ACT_RELU, "none": L.ACT_NONE, "sigmoid": L.ACT_SIGMOID}
        ops.conv_forward(xv, pc, yv, res=resv, act=acts[act], feed=feed, y_fp32=y_fp32, max_ctas=max_ctas,
                         n_tile=pc.n_tile)
        torch.cuda.synchronize()
        # reference from the packed (bf16, BN-folded) weights, fp32 math
        kd, kh, kw = k
        wq = pc.w[:cout, :kd * kh * kw * cin].float().reshape(cout, kd, kh, kw, cin)[..., :cin_real].permute(0, 4, 1, 2, 3)
        pad6 = (pad_f[2], pad_b[2], pad_f[1], pad_b[1], pad_f[0], pad_b[0])
        ref = conv_ref(bf(x), wq.contiguous(), pc.bias[:cout], stride, pad6, bf(res_t) if use_res else None, act)
        got = yv.to_ncdhw() if not y_fp32 else yv.interior().permute(0, 4, 1, 2, 3)
        ok = report(name, got, ref, extra=f"feed={'flat' if (feed == L.FEED_FLAT_TMA) else ('gather' if feed == L.FEED_GATHER else 'auto')} n_tile={pc.n_tile} M={N * od * oh * ow}")
        # halo check
        if sum(out_halo) > 0 and (feed != L.FEED_GATHER):
            full = yb.buf[..., out_coff:out_coff + cout].float().clone()
            full[:, out_halo[0]:out_halo[0] + od, out_halo[1]:out_halo[1] + oh, out_halo[2]:out_halo[2] + ow] = 0
            hz = full.abs().max().item()
            okh = hz == 0.0
            RESULTS.append((name + ":halo", okh))
            print(f"[{'PASS' if okh else 'FAIL'}] {name}: halo max |v| = {hz}")
        if ld_out != cout:
            other = torch.ones(ld_out, dtype=torch.bool)
            other[out_coff:out_coff + cout] = False
            leak = (yb.interior()[..., other.to(DEV)].float() - 3.0).abs().max().item()
            okl = leak == 0.0
            RESULTS.append((name + ":slice", okl))
            print(f"[{'PASS' if okl else 'FAIL'}] {name}: writes outside the channel slice: {leak}")
        return ok
    except Exception:
        RESULTS.append((name, False))
        print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
        return False


def group_flat():
    FL = L.FEED_FLAT_TMA
    run_conv_case("F1 2d 64->64 3x3", 2, (1, 12, 20), 64, 64, (1, 3, 3), halo=(0, 1, 1), feed=FL)
    run_conv_case("F1b 2d 64->64 3x3 1 cta", 2, (1, 12, 20), 64, 64, (1, 3, 3), halo=(0, 1, 1), feed=FL, max_ctas=1)
    run_conv_case("F2 2d 128->128 3x3", 3, (1, 16, 16), 128, 128, (1, 3, 3), halo=(0, 1, 1), feed=FL)
    run_conv_case("F3 2d 64->512 3x3 (2 n-tiles)", 2, (1, 12, 20), 64, 512, (1, 3, 3), halo=(0, 1, 1), feed=FL)
    run_conv_case("F4 2d 64->32 1x1", 2, (1, 12, 20), 64, 32, (1, 1, 1), halo=(0, 1, 1), feed=FL)
    run_conv_case("F5 2d slices+res", 2, (1, 14, 14), 64, 128, (1, 3, 3), halo=(0, 1, 1), feed=FL, use_res=True,
                  in_ld=192, in_coff=64, out_ld=256, out_coff=128)
    run_conv_case("F6 2d 256->256 56x56 8 ctas", 4, (1, 56, 56), 256, 256, (1, 3, 3), halo=(0, 1, 1), feed=FL, max_ctas=8)
    run_conv_case("F7 3d 64->64 3x3x3", 1, (4, 6, 6), 64, 64, (3, 3, 3), halo=(1, 1, 1), feed=FL)
    run_conv_case("F8 2d 512->512 14x14", 16, (1, 14, 14), 512, 512, (1, 3, 3), halo=(0, 1, 1), feed=FL)
    run_conv_case("F9 2d 1024->512 28x28 no-bn sigmoid", 2, (1, 28, 28), 1024, 512, (1, 3, 3), halo=(0, 1, 1), feed=FL,
                  act="sigmoid", bn=False)
    run_conv_case("F10 2d 64->16 3x3 (n_tile 16)", 2, (1, 12, 20), 64, 16, (1, 3, 3), halo=(0, 1, 1), feed=FL)


def run_multi_case(name, N, dhw, cin, couts):
    """b0 / b1a / b2a of an InceptionModule as ONE GEMM with three destinations (tedspad_conv y / y2 / y3)."""
    try:
        g = torch.Generator(device="cpu").manual_seed(cin + sum(couts))
        D, H, W = dhw
        x = torch.randn(N, cin, D, H, W, generator=g).to(DEV)
        xc = ops.CLTensor.from_ncdhw(x)
        pcs, refs = [], []
        for co in couts:
            w = (torch.randn(co, cin, 1, 1, 1, generator=g) / cin ** 0.5).to(DEV)
            bnp = ((torch.rand(co, generator=g) + 0.5).to(DEV), (torch.rand(co, generator=g) - 0.5).to(DEV),
                   (torch.rand(co, generator=g) - 0.5).to(DEV), (torch.rand(co, generator=g) + 0.5).to(DEV), 1e-3)
            pc = ops.PackedConv(w, None, bnp, device=DEV)
            pcs.append(pc)
            wq = pc.w[:co, :cin].float().reshape(co, cin, 1, 1, 1)
            refs.append(conv_ref(bf(x), wq, pc.bias[:co], (1, 1, 1), (0, 0, 0, 0, 0, 0), None, "relu"))
        pcm = ops.PackedConv.concat(pcs)
        # destinations: a slice of a wide buffer, a 64-channel padded scratch and a plain tensor
        big = ops.CLTensor(N, D, H, W, couts[0] + 40, device=DEV)
        big.buf.fill_(3.0)
        t1 = ops.CLTensor(N, D, H, W, max(128, couts[1] + 16), device=DEV)
        t1.buf.fill_(3.0)
        t2 = ops.CLTensor(N, D, H, W, couts[2], device=DEV)
        ys = [big.slice(8, couts[0]), t1.slice(0, couts[1]), t2]
        ops.conv_forward(xc, pcm, ys)
        torch.cuda.synchronize()
        for i, (yv, ref) in enumerate(zip(ys, refs)):
            report(f"{name}: destination {i} ({couts[i]} ch)", yv.to_ncdhw(), ref)
        untouched = bool((big.buf[..., :8] == 3.0).all() and (big.buf[..., 8 + couts[0]:] == 3.0).all() and
                         (t1.buf[..., couts[1]:] == 3.0).all())
        RESULTS.append((name + ":slices", untouched))
        print(f"[{'PASS' if untouched else 'FAIL'}] {name}: nothing written outside the destination slices")
    except Exception:
        RESULTS.append((name, False))
        print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_multi():
    run_multi_case("M1 Mixed_3b heads 192 -> 64|96|16 (FLAT)", 2, (4, 14, 14), 192, (64, 96, 16))
    run_multi_case("M2 Mixed_4b heads 480 -> 192|96|16 (GATHER)", 2, (4, 14, 14), 480, (192, 96, 16))
    run_multi_case("M3 Mixed_4c heads 512 -> 160|112|24", 2, (2, 7, 7), 512, (160, 112, 24))
    run_multi_case("M4 Mixed_5c heads 832 -> 384|192|48 (3 N tiles)", 3, (2, 7, 7), 832, (384, 192, 48))


def group_gather():
    G = L.FEED_GATHER
    run_conv_case("G1 2d 3(8)->64 3x3", 2, (1, 20, 20), 8, 64, (1, 3, 3), feed=G, cin_real=3)
    run_conv_case("G1b same conv as F1 via gather", 2, (1, 12, 20), 64, 64, (1, 3, 3), feed=G)
    run_conv_case("G1c gather into haloed out", 2, (1, 12, 20), 64, 64, (1, 3, 3), halo=(0, 1, 1), feed=G)
    run_conv_case("G2 3d stem 7x7x7 s2 TF-SAME", 1, (8, 20, 20), 8, 64, (7, 7, 7), stride=(2, 2, 2), pad_f=(2, 2, 2),
                  pad_b=(3, 3, 3), feed=G, cin_real=3)
    run_conv_case("G3 (1,3,3) s(1,2,2) odd", 2, (2, 11, 11), 128, 128, (1, 3, 3), stride=(1, 2, 2), pad_f=(0, 1, 1), feed=G)
    run_conv_case("G4 1x1x1 s(1,2,2) 256->512", 2, (2, 11, 11), 256, 512, (1, 1, 1), stride=(1, 2, 2), pad_f=(0, 0, 0), feed=G)
    run_conv_case("G5 3x3x3 16->32", 2, (4, 7, 7), 16, 32, (3, 3, 3), feed=G)
    run_conv_case("G5b 3x3x3 24->48 (Cout 48 -> n_tile 48)", 2, (4, 7, 7), 24, 48, (3, 3, 3), feed=G)
    run_conv_case("G5c 1x1x1 192->24 (n_tile 32)", 2, (4, 7, 7), 192, 24, (1, 1, 1), feed=G)
    run_conv_case("G6 linear 512->102 fp32", 3, (1, 1, 1), 512, 102, (1, 1, 1), feed=G, y_fp32=True, act="none", bn=False)
    run_conv_case("G7 (3,1,1) res relu", 2, (4, 7, 7), 64, 256, (3, 1, 1), pad_f=(1, 0, 0), feed=G, use_res=True)
    run_conv_case("G8 r3d 3x3x3 s2 64->128", 2, (8, 14, 14), 64, 128, (3, 3, 3), stride=(2, 2, 2), pad_f=(1, 1, 1), feed=G)
    run_conv_case("G9 many tiles 4 ctas", 4, (4, 28, 28), 64, 192, (3, 3, 3), feed=G, max_ctas=4)
    run_conv_case("G10 slice out + 832 in", 2, (2, 7, 7), 832, 384, (1, 1, 1), feed=G, out_ld=1024, out_coff=0)
    run_conv_case("G11 auto feed picks flat", 2, (1, 12, 20), 64, 64, (1, 3, 3), halo=(0, 1, 1), feed=L.FEED_AUTO)


def group_ops():
    g = torch.Generator(device="cpu").manual_seed(1)
    # max pools
    cases = [
        ("maxpool2d 2x2", (4, 64, 1, 12, 20), (1, 2, 2), (1, 2, 2), (0, 0, 0), (0, 0, 0), False),
        ("maxpool (1,3,3)s(1,2,2) SAME", (2, 64, 4, 14, 14), (1, 3, 3), (1, 2, 2), (0, 0, 0), (0, 1, 1), True),
        ("maxpool 3s2 SAME", (2, 32, 8, 14, 14), (3, 3, 3), (2, 2, 2), (0, 0, 0), (1, 1, 1), True),
        ("maxpool 3s1 SAME", (2, 32, 4, 7, 7), (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True),
        ("maxpool 3s1 SAME odd D/H/W", (3, 64, 3, 5, 9), (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True),
        ("maxpool 3s1 -inf pad", (2, 72, 2, 6, 4), (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), False),
        ("maxpool 3s1 SAME D=1", (2, 32, 1, 4, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True),
        ("maxpool 2s2", (2, 32, 4, 14, 14), (2, 2, 2), (2, 2, 2), (0, 0, 0), (0, 0, 0), True),
        ("maxpool (2,3,3)s2 p0 floor", (2, 64, 8, 23, 23), (2, 3, 3), (2, 2, 2), (0, 0, 0), (0, 0, 0), False),
        ("maxpool (2,1,1)", (2, 256, 4, 9, 9), (2, 1, 1), (2, 1, 1), (0, 0, 0), (0, 0, 0), False),
    ]
    for name, shp, k, s, pf, pb, zp in cases:
        try:
            x = torch.randn(*shp, generator=g).to(DEV)
            if zp:
                x = x - 1.0  # mostly negative so zero padding matters
            xc = ops.CLTensor.from_ncdhw(x, halo=(0, 1, 1))
            xpad = F.pad(bf(x), (pf[2], pb[2], pf[1], pb[1], pf[0], pb[0]), value=0.0 if zp else float("-inf"))
            ref = F.max_pool3d(xpad, k, s)
            yc = ops.CLTensor(shp[0], ref.shape[2], ref.shape[3], ref.shape[4], shp[1], device=DEV)
            ops.maxpool(xc, yc, k, s, pf, zero_pad=zp)
            report(name, yc.to_ncdhw(), ref, tol_rel=0)
        except Exception:
            RESULTS.append((name, False))
            print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # upsample x2 align_corners into a concat slice
    for name, (n, c, h, w), (th, tw) in [("upsample 7->14", (3, 64, 7, 7), (14, 14)),
                                          ("upsample 6x10 -> 13x21 (F.pad)", (2, 128, 6, 10), (13, 21))]:
        try:
            x = torch.randn(n, c, h, w, generator=g).to(DEV)
            xc = ops.CLTensor.from_ncdhw(x, halo=(0, 1, 1))
            cat = ops.CLTensor(n, 1, th, tw, 2 * c, (0, 1, 1), device=DEV)
            cat.buf.fill_(5.0)
            ops.upsample2x(xc, cat.slice(c, c))
            up = F.interpolate(bf(x), scale_factor=2, mode="bilinear", align_corners=True)
            dy, dx = th - up.shape[2], tw - up.shape[3]
            ref = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2]).unsqueeze(2)
            report(name, cat.slice(c, c).to_ncdhw(), ref, tol_rel=1e-2)
            okk = bool((cat.slice(0, c).interior().float() == 5.0).all())
            RESULTS.append((name + ":slice", okk))
            print(f"[{'PASS' if okk else 'FAIL'}] {name}: other slice untouched")
        except Exception:
            RESULTS.append((name, False))
            print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # nearest x2 into a concat slice (smp DecoderBlock.forward of the UNet++ anonymizer)
    for name, (n, c, h, w), (ld, coff) in [("nearest x2 7x9 -> slice of 192", (3, 64, 7, 9), (192, 64)),
                                           ("nearest x2 14x14 whole buffer", (2, 256, 14, 14), (256, 0))]:
        try:
            x = torch.randn(n, c, h, w, generator=g).to(DEV)
            src = ops.CLTensor(n, 1, h, w, c + 64, (0, 1, 1), device=DEV)   # source is itself a channel slice
            src.slice(64, c).interior()[:, 0] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
            cat = ops.CLTensor(n, 1, 2 * h, 2 * w, ld, (0, 1, 1), device=DEV)
            cat.buf.fill_(5.0)
            ops.upsample2x_nearest(src.slice(64, c), cat.slice(coff, c))
            ref = F.interpolate(bf(x), scale_factor=2, mode="nearest").unsqueeze(2)
            report(name, cat.slice(coff, c).to_ncdhw(), ref, tol_rel=0)
            full = cat.buf.float().clone()
            full[:, :, 1:1 + 2 * h, 1:1 + 2 * w, coff:coff + c] = 5.0
            okk = bool((full == 5.0).all())
            RESULTS.append((name + ":rest", okk))
            print(f"[{'PASS' if okk else 'FAIL'}] {name}: halo and other channels untouched")
        except Exception:
            RESULTS.append((name, False))
            print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # channels-last head output -> encoder clip through the raw-reshape glue (+ fp32 frames)
    try:
        B, T, H, W = 2, 16, 10, 12
        fr = (torch.randn(B * T, 3, H, W, generator=g) * 2).to(DEV)
        ref = bf(fr).reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)
        for cpad, halo in ((4, (0, 0, 0)), (8, (0, 0, 0))):
            xc = ops.CLTensor.from_ncdhw(fr, halo=(0, 1, 1))       # [B*T,1,H,W,8], channels 3..7 zero
            xc.interior()[..., 3:] = 7.0                            # junk in the pad channels must not leak
            enc = ops.CLTensor(B, T, H, W, cpad, device=DEV)
            enc.buf.fill_(5.0)
            out = torch.full((B * T, 3, H, W), 9.0, device=DEV)
            ops.frames_to_clip(xc, enc, T, out)
            report(f"frames_to_clip C={cpad}", enc.to_ncdhw()[:, :3], ref, tol_rel=0)
            report(f"frames_to_clip C={cpad} fp32 frames", out, bf(fr), tol_rel=0)
            # the same frames handed over in space-to-depth form [B*T,1,H/2,W/2,16] (channel (2a+b)*3 + c)
            s2d = fr.reshape(B * T, 3, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B * T, H // 2, W // 2, 12)
            xs = ops.CLTensor(B * T, 1, H // 2, W // 2, 16, (0, 1, 1), device=DEV)
            xs.buf.fill_(7.0)
            xs.interior()[:, 0, :, :, :12] = s2d.to(torch.bfloat16)
            enc.buf.fill_(5.0)
            out.fill_(9.0)
            ops.frames_to_clip(xs, enc, T, out, s2d=True)
            report(f"frames_to_clip s2d C={cpad}", enc.to_ncdhw()[:, :3], ref, tol_rel=0)
            report(f"frames_to_clip s2d C={cpad} fp32 frames", out, bf(fr), tol_rel=0)
            okz = bool((enc.interior()[..., 3:] == 0).all())
            RESULTS.append((f"frames_to_clip C={cpad} pad", okz))
            print(f"[{'PASS' if okz else 'FAIL'}] frames_to_clip C={cpad}: pad channels zero")
    except Exception:
        RESULTS.append(("frames_to_clip", False))
        print(f"[FAIL] frames_to_clip: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # outconv + sigmoid + scatter
    try:
        B, T, H, W, Cc = 2, 16, 10, 12, 64
        x = torch.randn(B * T, Cc, H, W, generator=g).to(DEV)
        w = (torch.randn(3, Cc, generator=g) / 8).to(DEV)
        b = torch.randn(3, generator=g).to(DEV)
        xc = ops.CLTensor.from_ncdhw(x, halo=(0, 1, 1))
        enc = ops.CLTensor(B, T, H, W, 8, device=DEV)
        enc.buf.zero_()
        fr = torch.empty(B * T, 3, H, W, device=DEV)
        ops.outconv_sigmoid(xc, w, b, enc, T, fr)
        ref_fr = torch.sigmoid(F.conv2d(bf(x), w[:, :, None, None], b))
        report("outconv frames fp32", fr, ref_fr, tol_rel=1e-5)
        ref_enc = ref_fr.reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)  # dali_extraction.py:171-173 raw reshape
        report("outconv scatter (raw reshape glue)", enc.to_ncdhw()[:, :3], ref_enc, tol_rel=5e-3)
        okz = bool((enc.interior()[..., 3:] == 0).all())
        RESULTS.append(("outconv pad channels zero", okz))
        print(f"[{'PASS' if okz else 'FAIL'}] outconv pad channels stay zero")
    except Exception:
        RESULTS.append(("outconv", False))
        print(f"[FAIL] outconv: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # avgpool features
    try:
        x = torch.randn(3, 1024, 2, 7, 7, generator=g).to(DEV)
        xc = ops.CLTensor.from_ncdhw(x)
        out = ops.avgpool_features(xc, 2)
        ref = F.avg_pool3d(bf(x), (2, 7, 7), 1).reshape(3, 1024, 1).permute(0, 2, 1)
        report("avgpool (2,7,7)", out, ref, tol_rel=1e-5)
        x = torch.randn(2, 64, 4, 7, 7, generator=g).to(DEV)
        out = ops.avgpool_features(ops.CLTensor.from_ncdhw(x), 2)
        ref = F.avg_pool3d(bf(x), (2, 7, 7), 1).reshape(2, 64, 3).permute(0, 2, 1)
        report("avgpool sliding D=4", out, ref, tol_rel=1e-5)
        out = ops.avgpool_features(ops.CLTensor.from_ncdhw(x), 0)
        report("avgpool global", out, bf(x).mean(dim=(2, 3, 4)).unsqueeze(1), tol_rel=1e-5)
        x = torch.randn(5, 2048, 2, 7, 7, generator=g).to(DEV)       # I3Res50 head: sliced input, odd window count
        wide = ops.CLTensor(5, 2, 7, 7, 2048 + 64, device=DEV)
        wide.slice(64, 2048).interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
        out = ops.avgpool_features(wide.slice(64, 2048), 0)
        report("avgpool global 2048 from a channel slice", out, bf(x).mean(dim=(2, 3, 4)).unsqueeze(1), tol_rel=1e-5)
    except Exception:
        RESULTS.append(("avgpool", False))
        print(f"[FAIL] avgpool: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # row-wise L2 normalisation (mlp embedding head) and the linear helper
    try:
        x = torch.randn(5, 128, generator=g).to(DEV)
        x[3] = 0
        want = F.normalize(x, p=2, dim=1)
        got = ops.l2_normalize_rows(x.clone())
        report("l2_normalize_rows", got, want, tol_rel=1e-6)
        w = torch.randn(24, 40, generator=g).to(DEV)
        b = torch.randn(24, generator=g).to(DEV)
        pc = ops.PackedConv(w, b, None, device=DEV)
        from tedspad_b200.engine import _Buffers
        xin = torch.randn(7, 40, generator=g).to(DEV)
        y = ops.linear(xin, pc, (_Buffers(DEV), "lin"))
        report("linear 40->24 through the convolution kernel", y, F.linear(bf(xin), bf(w), b), tol_rel=1e-2)
    except Exception:
        RESULTS.append(("l2_normalize / linear", False))
        print(f"[FAIL] l2_normalize / linear: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # nchw -> channels-last
    try:
        x = torch.rand(4, 3, 20, 24, generator=g).to(DEV)
        y = ops.CLTensor(4, 1, 20, 24, 8, device=DEV)
        y.buf.fill_(9.0)
        ops.nchw_to_cl(x, y)
        report("nchw_to_cl C=3->8", y.to_ncdhw()[:, :3, 0], bf(x), tol_rel=0)
        x5 = torch.rand(2, 3, 4, 6, 10, generator=g).to(DEV)
        y5 = ops.CLTensor(2, 4, 6, 10, 8, device=DEV)
        ops.nchw_to_cl(x5, y5)
        report("nchw_to_cl 5-D", y5.to_ncdhw()[:, :3], bf(x5), tol_rel=0)
        okz = bool((y.interior()[..., 3:] == 0).all())
        RESULTS.append(("nchw_to_cl zero pad", okz))
    except Exception:
        RESULTS.append(("nchw_to_cl", False))
        print(f"[FAIL] nchw_to_cl: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_prep():
    import numpy as np
    import torchvision.transforms.functional as TF
    from PIL import Image
    g = torch.Generator(device="cpu").manual_seed(2)
    # DALI path: 240x320 -> crop 192x256 -> 224x224, antialias float
    try:
        Fr, Hs, Ws = 5, 240, 320
        frames = torch.randint(0, 256, (Fr, Hs, Ws, 3), generator=g, dtype=torch.uint8)
        ch, cw = int(Hs * 0.8), int(Ws * 0.8)
        top, left = int(round((Hs - ch) / 2.0)), int(round((Ws - cw) / 2.0))
        desc = torch.tensor([[0, top, left, 0], [3, top, left, 0], [-1, top, left, 0], [4, 0, 0, 0], [4, 0, 0, 1]],
                            dtype=torch.int32)
        y = ops.CLTensor(5, 1, 224, 224, 8, device=DEV)
        f32 = torch.empty(5, 3, 224, 224, device=DEV)
        ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), y, L.RESAMPLE_AA_FLOAT, f32)
        v = frames.permute(0, 3, 1, 2).float() / 255.0
        refs = []
        for s, t, l, fl in desc.tolist():
            if s < 0:
                refs.append(torch.zeros(3, 224, 224))
                continue
            img = v[s]
            if fl:
                img = img.flip(-1)
            img = img[:, t:t + ch, l:l + cw]
            refs.append(TF.resize(img, (224, 224), antialias=True))
        ref = torch.stack(refs).to(DEV)
        report("preprocess AA float fp32", f32, ref, tol_rel=3e-5)
        report("preprocess AA bf16 layout", y.to_ncdhw()[:, :3, 0], bf(ref), tol_rel=8e-3)
        # the same call without the fp32 copy takes the FAST instantiation (compile-time band, vector stores): its
        # pixels must equal the generic instantiation's bit for bit, for 8- and for 4-channel pixels
        for cpad in (8, 4):
            ya = ops.CLTensor(5, 1, 224, 224, cpad, device=DEV)
            yb = ops.CLTensor(5, 1, 224, 224, cpad, device=DEV)
            ya.buf.fill_(3.0); yb.buf.fill_(4.0)
            ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), ya, L.RESAMPLE_AA_FLOAT, f32)
            ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), yb, L.RESAMPLE_AA_FLOAT, None)
            okf = bool(torch.equal(ya.buf, yb.buf))
            RESULTS.append((f"preprocess fast == generic C={cpad}", okf))
            print(f"[{'PASS' if okf else 'FAIL'}] preprocess AA: FAST instantiation == generic, {cpad}-channel pixels")
        cc = TF.center_crop(v[0:1], (ch, cw))
        ok = torch.equal(cc, v[0:1, :, top:top + ch, left:left + cw])
        RESULTS.append(("center_crop offsets", ok))
        print(f"[{'PASS' if ok else 'FAIL'}] center_crop offset formula")
    except Exception:
        RESULTS.append(("preprocess AA", False))
        print(f"[FAIL] preprocess AA: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # larger down-scale (XD-Violence-shaped 360x640, 10-crop boxes incl. a flipped corner) into 4-channel pixels (the
    # UNet++ stem's input layout): the 5 / 8-tap instantiations of the kernel
    try:
        Fr, Hs, Ws = 3, 360, 640
        frames = torch.randint(0, 256, (Fr, Hs, Ws, 3), generator=g, dtype=torch.uint8)
        ch, cw = int(Hs * 0.8), int(Ws * 0.8)
        desc = torch.tensor([[0, 0, Ws - cw, 0], [2, Hs - ch, 0, 1], [1, 36, 64, 0]], dtype=torch.int32)
        for reso in ((224, 224), (160, 192)):      # 6 x 4 and 8 x 5 taps (anything beyond 8 taps per axis is refused)
            y = ops.CLTensor(3, 1, reso[0], reso[1], 4, device=DEV)
            y.buf.fill_(9.0)
            f32 = torch.empty(3, 3, reso[0], reso[1], device=DEV)
            ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), y, L.RESAMPLE_AA_FLOAT, f32)
            v = frames.permute(0, 3, 1, 2).float() / 255.0
            refs = []
            for s_, t, l, fl in desc.tolist():
                img = v[s_].flip(-1) if fl else v[s_]
                refs.append(TF.resize(img[:, t:t + ch, l:l + cw], reso, antialias=True))
            ref = torch.stack(refs).to(DEV)
            report(f"preprocess AA 360x640 -> {reso[0]} fp32 (down-scale taps)", f32, ref, tol_rel=3e-5)
            report(f"preprocess AA 360x640 -> {reso[0]} bf16, 4-channel pixels", y.to_ncdhw()[:, :3, 0], bf(ref), tol_rel=8e-3)
            okz = bool((y.interior()[..., 3] == 0).all())
            RESULTS.append((f"preprocess C=4 pad {reso[0]}", okz))
            print(f"[{'PASS' if okz else 'FAIL'}] preprocess 4-channel pixels: pad channel zero")
    except Exception:
        RESULTS.append(("preprocess AA downscale", False))
        print(f"[FAIL] preprocess AA downscale: EXCEPTION\n{traceback.format_exc()}", flush=True)
    # ShanghaiTech path: 480x856 -> crop 384x384 -> PIL bilinear uint8
    try:
        Fr, Hs, Ws = 2, 480, 856
        frames = torch.randint(0, 256, (Fr, Hs, Ws, 3), generator=g, dtype=torch.uint8)
        # add smooth content too
        yy, xx = torch.meshgrid(torch.arange(Hs), torch.arange(Ws), indexing="ij")
        frames[1] = torch.stack([(yy * 255 // Hs), (xx * 255 // Ws), ((yy + xx) % 256)], -1).to(torch.uint8)
        ch = cw = int(Hs * 0.8)
        top, left = int(round((Hs - ch) / 2.0)), int(round((Ws - cw) / 2.0))
        desc = torch.tensor([[0, top, left, 0], [1, top, left, 0]], dtype=torch.int32)
        y = ops.CLTensor(2, 1, 224, 224, 8, device=DEV)
        f32 = torch.empty(2, 3, 224, 224, device=DEV)
        ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), y, L.RESAMPLE_PIL_U8, f32)
        refs = []
        for i in range(2):
            im = TF.to_pil_image(frames[i].numpy())
            im = TF.center_crop(im, (ch, cw))
            im = TF.resize(im, (224, 224), antialias=True)
            refs.append(TF.to_tensor(im))
        ref = torch.stack(refs).to(DEV)
        d = ((f32 - ref).abs() * 255).round()
        print(f"       PIL emulation: exact={float((d == 0).float().mean()):.6f} max_lsb={int(d.max())}")
        report("preprocess PIL u8 (bit-exact)", f32, ref, tol_rel=0)
        yb = ops.CLTensor(2, 1, 224, 224, 8, device=DEV)
        ops.preprocess(frames.to(DEV), desc.to(DEV), (ch, cw), yb, L.RESAMPLE_PIL_U8, None)     # FAST instantiation
        okf = bool(torch.equal(y.buf, yb.buf))
        RESULTS.append(("preprocess PIL fast == generic", okf))
        print(f"[{'PASS' if okf else 'FAIL'}] preprocess PIL: FAST instantiation == generic")
    except Exception:
        RESULTS.append(("preprocess PIL", False))
        print(f"[FAIL] preprocess PIL: EXCEPTION\n{traceback.format_exc()}", flush=True)


def time_conv(name, N, hw, cin, cout, k=(1, 3, 3), iters=10):
    try:
        H, W = hw
        halo = (0, 1, 1)
        x = ops.CLTensor(N, 1, H, W, cin, halo, device=DEV)
        x.interior().normal_()
        wt = torch.randn(cout, cin, *k, device=DEV) / (cin * 9) ** 0.5
        pc = ops.PackedConv(wt, None, None, pad_front=tuple(kk // 2 for kk in k), cin_pad=cin, device=DEV)
        y = ops.CLTensor(N, 1, H, W, cout, halo, device=DEV)
        for _ in range(3):
            ops.conv_forward(x, pc, y, feed=L.FEED_FLAT_TMA, n_tile=pc.n_tile)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv_forward(x, pc, y, feed=L.FEED_FLAT_TMA, n_tile=pc.n_tile)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * N * H * W * cin * cout * k[0] * k[1] * k[2]
        print(f"[PERF] {name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)  n_tile={pc.n_tile}", flush=True)
    except Exception:
        print(f"[FAIL] perf {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_perf():
    time_conv("unet 64->64 @224 x16", 16, (224, 224), 64, 64)
    time_conv("unet 128->64 @224 x16", 16, (224, 224), 128, 64)
    time_conv("unet 128->128 @112 x16", 16, (112, 112), 128, 128)
    time_conv("unet 256->256 @56 x16", 16, (56, 56), 256, 256)
    time_conv("unet 512->512 @28 x16", 16, (28, 28), 512, 512)
    time_conv("unet 1024->512 @28 x16", 16, (28, 28), 1024, 512)
    time_conv("unet 512->512 @14 x16", 16, (14, 14), 512, 512)
    time_conv("unet 256->256 @56 x64", 64, (56, 56), 256, 256)
    time_conv("unet 512->512 @28 x64", 64, (28, 28), 512, 512)


# ------------------------------------------------------------------------------------------ SLAB feed
def run_slab_case(name, kind, N, dhw, cin_real, cin_buf, cout, k, stride=(1, 1, 1), pad_f=(0, 1, 1), pad_b=None, halo=(0, 0, 0),
                  tm=0, in_ld=None, in_coff=0, out_ld=None, out_coff=0, out_halo=None, pool=False, outconv=False,
                  max_ctas=0, seed=0, res=False):
    import numpy as np
    import _slabsim as S
    try:
        g = torch.Generator(device="cpu").manual_seed(seed)
        D, H, W = dhw
        pad_b = pad_b if pad_b is not None else pad_f
        x = torch.randn(N, cin_real, D, H, W, generator=g).to(DEV)
        w = (torch.randn(cout, cin_real, *k, generator=g) / (cin_real * k[0] * k[1] * k[2]) ** 0.5).to(DEV)
        b = (torch.rand(cout, generator=g) - 0.5).to(DEV)
        bnp = ((torch.rand(cout, generator=g) + 0.5).to(DEV), (torch.rand(cout, generator=g) - 0.5).to(DEV),
               (torch.rand(cout, generator=g) - 0.5).to(DEV), (torch.rand(cout, generator=g) + 0.5).to(DEV), 1e-3)
        std_cin = cin_buf if kind in (L.SLAB_3X3, L.SLAB_3X3_STREAM, L.SLAB_3X3_PAIR, L.SLAB_3X3_STREAM_PAIR, L.SLAB_3X3_KX_PAIR) else 8
        pc = ops.PackedConv(w, b, bnp, stride=stride, pad_front=pad_f, cin_pad=std_cin, device=DEV, n_align=32)
        psc = ops.PackedSlabConv(pc, kind)
        if kind not in (L.SLAB_3X3_STREAM, L.SLAB_3X3_STREAM_PAIR):
            # the CUDA pack kernel against the numpy restatement, bit for bit
            img_ref = S.pack_image(kind, S.bf16_bits(pc.w), pc.cout_pad, pc.k_pad, pc.cin_pad, pc.k, pad_f[2])
            okp = bool(np.array_equal(S.bf16_bits(psc.image), img_ref))
            RESULTS.append((name + ":pack", okp))
            print(f"[{'PASS' if okp else 'FAIL'}] {name}: weight image == numpy packer ({img_ref.size * 2} B)")
        od, oh, ow = pc.out_extent((D, H, W), pad_b)
        ld_in = in_ld or cin_buf
        xb = ops.CLTensor(N, D, H, W, ld_in, halo, device=DEV)
        if ld_in != cin_buf:
            xb.interior()[...] = 7.0
        xv = xb.slice(in_coff, cin_buf)
        xv.interior()[..., :cin_real] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
        if cin_real < cin_buf:
            xv.interior()[..., cin_real:] = 0
        out_halo = out_halo if out_halo is not None else (halo if (od, oh, ow) == (D, H, W) else (0, 0, 0))
        ld_out = out_ld or cout
        yb = ops.CLTensor(N, od, oh, ow, ld_out, out_halo, device=DEV)
        yb.buf.fill_(3.0)
        yv = yb.slice(out_coff, cout)
        pv = None
        if pool:
            pb = ops.CLTensor(N, 1, oh // 2, ow // 2, cout, (0, 1, 1), device=DEV)
            pb.buf.fill_(3.0)
            pv = pb
        oc = None
        if outconv:
            ocw = (torch.randn(3, cout, generator=g) / 8).to(DEV)
            ocb = torch.randn(3, generator=g).to(DEV)
            planes = torch.full((N, 3, oh, ow), 9.0, device=DEV, dtype=torch.bfloat16)
            frames = torch.full((N, 3, oh, ow), 9.0, device=DEV)
            Tc = 2 if N % 2 == 0 else 1
            clip = ops.CLTensor(N // Tc, Tc, oh, ow, 4, device=DEV)
            clip.buf.fill_(9.0)
            oc = (ocw, ocb, planes, frames, clip, Tc)
        rv, rres = None, None
        if res:   # bf16 residual living in a channel slice of a wider buffer (Bottleneck tail: relu(conv + bias + res))
            rres = torch.randn(N, cout, od, oh, ow, generator=g).to(DEV)
            rb = ops.CLTensor(N, od, oh, ow, cout + 16, out_halo, device=DEV)
            rb.buf.fill_(5.0)
            rv = rb.slice(8, cout)
            rv.interior()[...] = rres.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
        ops.conv_slab_forward(xv, psc, yv, pool=pv, outconv=oc, tm=tm, max_ctas=max_ctas, res=rv)
        torch.cuda.synchronize()
        kd, kh, kw = k
        wq = pc.w[:cout, :kd * kh * kw * pc.cin_pad].float().reshape(cout, kd, kh, kw, pc.cin_pad)[..., :cin_real].permute(0, 4, 1, 2, 3)
        pad6 = (pad_f[2], pad_b[2], pad_f[1], pad_b[1], pad_f[0], pad_b[0])
        ref = conv_ref(bf(x), wq.contiguous(), pc.bias[:cout], stride, pad6, bf(rres) if res else None, "relu")
        plan = psc.plan(xv, yv, tm=tm)
        ok = report(name, yv.to_ncdhw(), ref, extra=f"tm={plan.tm} stages={plan.stages} b_stages={plan.b_stages} k_stages={plan.k_stages} "
                    f"n_tile={plan.n_tile}x{plan.num_n_tiles} tiles={plan.total_tiles} smem={plan.smem_bytes}")
        if sum(out_halo) > 0:
            full = yb.buf[..., out_coff:out_coff + cout].float().clone()
            full[:, out_halo[0]:out_halo[0] + od, out_halo[1]:out_halo[1] + oh, out_halo[2]:out_halo[2] + ow] = 3.0
            okh = bool((full == 3.0).all())
            RESULTS.append((name + ":halo", okh))
            print(f"[{'PASS' if okh else 'FAIL'}] {name}: halo untouched")
        if ld_out != cout:
            other = torch.ones(ld_out, dtype=torch.bool)
            other[out_coff:out_coff + cout] = False
            leak = (yb.interior()[..., other.to(DEV)].float() - 3.0).abs().max().item()
            RESULTS.append((name + ":slice", leak == 0.0))
            print(f"[{'PASS' if leak == 0.0 else 'FAIL'}] {name}: writes outside the channel slice: {leak}")
        if pool:
            pref = F.max_pool2d(yv.to_ncdhw()[:, :, 0], 2).unsqueeze(2)   # pooled from the kernel's own bf16 output
            report(name + ":pool", pv.to_ncdhw(), pref, tol_rel=0)
            fullp = pb.buf.float().clone()
            fullp[:, :, 1:1 + oh // 2, 1:1 + ow // 2] = 3.0
            okh = bool((fullp == 3.0).all())
            RESULTS.append((name + ":pool halo", okh))
            print(f"[{'PASS' if okh else 'FAIL'}] {name}: pool halo untouched")
        if outconv:
            fr_ref = torch.sigmoid(F.conv2d(ref[:, :, 0], ocw[:, :, None, None], ocb))
            report(name + ":outconv frames", frames, fr_ref, tol_rel=2e-3)
            report(name + ":outconv planes", planes.float(), fr_ref, tol_rel=6e-3)
            # the same images through the raw-reshape glue (dali_extraction.py:171-173), straight into the encoder clip
            enc_ref = fr_ref.reshape(N // Tc, Tc, 3, oh, ow).reshape(N // Tc, 3, Tc, oh, ow)
            report(name + ":outconv clip (glue)", clip.to_ncdhw()[:, :3], enc_ref, tol_rel=6e-3)
            okc = bool(torch.equal(clip.to_ncdhw()[:, :3].to(torch.bfloat16).reshape(N // Tc, 3 * Tc, oh, ow),
                                   planes.reshape(N // Tc, 3 * Tc, oh, ow))) and bool((clip.interior()[..., 3] == 9.0).all())
            RESULTS.append((name + ":clip==planes", okc))
            print(f"[{'PASS' if okc else 'FAIL'}] {name}: clip channels == planes bit for bit, pad channel untouched")
        return ok
    except Exception:
        RESULTS.append((name, False))
        print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)
        return False


def group_slab3():
    K = L.SLAB_3X3
    run_slab_case("S1 64->64 20x24 tm2", K, 2, (1, 20, 24), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), tm=2)
    run_slab_case("S2 64->64 20x24 tm1 1cta", K, 2, (1, 20, 24), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), tm=1, max_ctas=1)
    run_slab_case("S3 128->64 slices auto", K, 3, (1, 32, 32), 128, 128, 64, (1, 3, 3), halo=(0, 1, 1), in_ld=192, in_coff=64,
                  out_ld=128, out_coff=64)
    run_slab_case("S4 64->128 28x28", K, 2, (1, 28, 28), 64, 64, 128, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("S5 64->64 pool fused 32x48", K, 3, (1, 32, 48), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), pool=True,
                  out_ld=128, out_coff=0)
    run_slab_case("S6 64->64 outconv fused", K, 4, (1, 32, 32), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), outconv=True)
    run_slab_case("S7 64->64 112x112 x8 many tiles", K, 8, (1, 112, 112), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("S8 64->32 odd 19x21", K, 2, (1, 19, 21), 64, 64, 32, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("S9 (1,3,3) D=3 64->64", K, 2, (3, 16, 16), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    # UNet++ tail: 32 real input channels stored as 64, 8 output channels (the 3x3 head, 3 real + 5 zero rows)
    run_slab_case("S10 32(64)->8 head 40x40 x3", K, 3, (1, 40, 40), 32, 64, 8, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("S11 64->32 into a 64-wide buffer", K, 2, (1, 32, 32), 64, 64, 32, (1, 3, 3), halo=(0, 1, 1), out_ld=64, out_coff=0)
    run_slab_case("S12 64->64 + residual (BasicBlock tail) 56x56 x3 stacked", K, 3, (1, 56, 56), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), res=True)


def group_slabpair():
    """CTA-pair kind (cta_group::2): same cases as the single-CTA 3X3 kind wherever the tile count is even."""
    K = L.SLAB_3X3_PAIR
    run_slab_case("P1 64->64 20x24 pair", K, 2, (1, 20, 24), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("P2 64->64 20x24 pair 2 ctas", K, 2, (1, 20, 24), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), max_ctas=2)
    run_slab_case("P3 128->64 slices pair", K, 4, (1, 32, 32), 128, 128, 64, (1, 3, 3), halo=(0, 1, 1), in_ld=192, in_coff=64,
                  out_ld=128, out_coff=64)
    run_slab_case("P4 64->64 pool fused 32x48 pair", K, 3, (1, 32, 48), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), pool=True,
                  out_ld=128, out_coff=0)
    run_slab_case("P5 64->64 outconv fused pair", K, 4, (1, 32, 32), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), outconv=True)
    run_slab_case("P6 64->64 112x112 x8 many tiles pair", K, 8, (1, 112, 112), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("P7 64->64 odd 19x21 x2 pair", K, 2, (1, 19, 21), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("P8 64->64 odd tile count -> single-CTA fallback", K, 1, (1, 16, 48), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("P9 64->128 28x28 pair (N=128)", K, 2, (1, 28, 28), 64, 64, 128, (1, 3, 3), halo=(0, 1, 1))


def group_slabkx():
    """KX kind: the three taps of a filter row share one A-operand fetch (N = 3 * Cout_pad), 8 x 16 tiles, 14 output columns."""
    K = L.SLAB_3X3_KX_PAIR
    run_slab_case("K1 64->64 16x28 kx", K, 2, (1, 16, 28), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K2 64->64 20x24 kx (ragged rows / columns)", K, 2, (1, 20, 24), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K3 128->64 slices kx", K, 4, (1, 32, 32), 128, 128, 64, (1, 3, 3), halo=(0, 1, 1), in_ld=192, in_coff=64,
                  out_ld=128, out_coff=64)
    run_slab_case("K4 64->64 112x112 x8 many tiles kx", K, 8, (1, 112, 112), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K5 64->64 odd 19x21 x2 kx (stacked rows)", K, 2, (1, 19, 21), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K6 128->16 head 40x42 x4 kx (Cout_pad 32)", K, 4, (1, 40, 42), 128, 128, 16, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K6b 128->16 head 40x42 x3 kx (odd tile count -> fallback)", K, 3, (1, 40, 42), 128, 128, 16, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K6c 64->8 kx (Cout 8 of 32), no activation", K, 2, (1, 24, 28), 64, 64, 8, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("K7 64->64 pool fused -> fallback", K, 3, (1, 32, 48), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), pool=True,
                  out_ld=128, out_coff=0)
    run_slab_case("K8 64->64 + residual -> fallback", K, 3, (1, 56, 56), 64, 64, 64, (1, 3, 3), halo=(0, 1, 1), res=True)
    run_slab_case("K9 64->40 kx (Cout 40 of 64)", K, 2, (1, 24, 30), 64, 64, 40, (1, 3, 3), halo=(0, 1, 1))


    # the head's raw-reshape glue fused into the KX epilogue == frames_to_clip(s2d) of the stored head tensor, bit for bit
    try:
        g = torch.Generator(device="cpu").manual_seed(9)
        N, T, H, W = 8, 4, 24, 28
        x = ops.CLTensor(N, 1, H, W, 128, (0, 1, 1), device=DEV)
        x.interior().normal_(generator=None)
        wt = torch.randn(16, 128, 1, 3, 3, generator=g).to(DEV) / 34.0
        wt[12:] = 0
        pc = ops.PackedConv(wt, torch.randn(16, generator=g).to(DEV) * 0.1, None, pad_front=(0, 1, 1), cin_pad=128, device=DEV, n_align=32)
        psc = ops.PackedSlabConv(pc, K)
        okk = ops.slab_runs_kx(x, psc)
        y = ops.CLTensor(N, 1, H, W, 16, (0, 1, 1), device=DEV)
        clips = []
        for fused in (True, False):
            clip = ops.CLTensor(N // T, T, 2 * H, 2 * W, 4, device=DEV)
            clip.buf.zero_()
            if fused:
                ops.conv_slab_forward(x, psc, y, act=L.ACT_NONE, s2d_clip=(clip, T))
            else:
                ops.frames_to_clip(y, clip, T, None, s2d=True)
            clips.append(clip.buf.clone())
        clip_only = ops.CLTensor(N // T, T, 2 * H, 2 * W, 4, device=DEV)
        clip_only.buf.zero_()
        ops.conv_slab_forward(x, psc, None, act=L.ACT_NONE, s2d_clip=(clip_only, T))   # no head tensor at all
        torch.cuda.synchronize()
        ok = okk and bool(torch.equal(clips[0], clips[1])) and bool(torch.equal(clip_only.buf, clips[1])) and float(clips[1].abs().sum()) > 0
        RESULTS.append(("K10 kx fused s2d clip glue", ok))
        print(f"[{'PASS' if ok else 'FAIL'}] K10 kx head: fused clip glue == frames_to_clip(s2d) bit for bit (with and without the head tensor), kx={okk}")
    except Exception:
        RESULTS.append(("K10 kx fused s2d clip glue", False))
        print(f"[FAIL] K10: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_kxperf():
    for K, nm in ((L.SLAB_3X3_PAIR, "pair"), (L.SLAB_3X3_KX_PAIR, "kx")):
        time_slab(f"64->64 @224 x128 {nm}", K, 128, (1, 224, 224), 64, 64, (1, 3, 3))
        time_slab(f"128->64 @224 x128 {nm}", K, 128, (1, 224, 224), 128, 64, (1, 3, 3))
        time_slab(f"128->64 @112 x128 {nm}", K, 128, (1, 112, 112), 128, 64, (1, 3, 3))
        time_slab(f"64->64 @112 x512 {nm}", K, 512, (1, 112, 112), 64, 64, (1, 3, 3))
        time_slab(f"64->64 @56 x512 {nm}", K, 512, (1, 56, 56), 64, 64, (1, 3, 3))
    time_slab("128->16 head @112 x512 3x3", L.SLAB_3X3, 512, (1, 112, 112), 128, 16, (1, 3, 3))
    time_slab("128->16 head @112 x512 kx", L.SLAB_3X3_KX_PAIR, 512, (1, 112, 112), 128, 16, (1, 3, 3))


def group_slab1x1():
    """One spatial tap per K stage (1x1x1 and (3,1,1)) and the residual epilogue of the streaming kind."""
    K = L.SLAB_3X3_STREAM
    run_slab_case("T1 1x1x1 64->64 Conv3d_2b shape", K, 2, (4, 28, 28), 64, 64, 64, (1, 1, 1), pad_f=(0, 0, 0))
    run_slab_case("T2 1x1x1 256->64 bottleneck conv1 odd 55x55", K, 2, (2, 55, 55), 256, 256, 64, (1, 1, 1), pad_f=(0, 0, 0))
    run_slab_case("T3 (3,1,1) 256->64 temporal conv1", K, 2, (4, 23, 23), 256, 256, 64, (3, 1, 1), pad_f=(1, 0, 0))
    run_slab_case("T4 1x1x1 64->256 conv3 + residual", K, 2, (2, 55, 55), 64, 64, 256, (1, 1, 1), pad_f=(0, 0, 0), res=True)
    run_slab_case("T5 1x1x1 512->2048 conv3 + residual (8 N tiles)", K, 2, (2, 7, 7), 512, 512, 2048, (1, 1, 1), pad_f=(0, 0, 0), res=True)
    run_slab_case("T6 1x1x1 192->32 pool branch, slice out", K, 2, (4, 14, 14), 192, 192, 32, (1, 1, 1), pad_f=(0, 0, 0),
                  out_ld=256, out_coff=224)
    run_slab_case("T7 3x3x3 64->64 + residual (R3D BasicBlock conv2)", K, 2, (4, 28, 28), 64, 64, 64, (3, 3, 3), pad_f=(1, 1, 1), res=True)
    run_slab_case("T8 2-D 1x1 128->128 haloed (stacked rows)", K, 3, (1, 20, 24), 128, 128, 128, (1, 1, 1), pad_f=(0, 0, 0),
                  halo=(0, 1, 1))
    run_slab_case("T9 1x1x1 832->128 7x7", K, 4, (2, 7, 7), 832, 832, 128, (1, 1, 1), pad_f=(0, 0, 0))
    # strided 1x1x1 (down-sample projections): the TMA box walks the input with the convolution's stride
    run_slab_case("T10 2-D 1x1 s2 64->128 56x56 haloed (resnet18 layer2.0.downsample)", K, 3, (1, 56, 56), 64, 64, 128, (1, 1, 1),
                  stride=(1, 2, 2), pad_f=(0, 0, 0), halo=(0, 1, 1), out_halo=(0, 1, 1))
    run_slab_case("T11 2-D 1x1 s2 128->256 odd 23x21, slice in / out", K, 2, (1, 23, 21), 128, 128, 256, (1, 1, 1), stride=(1, 2, 2),
                  pad_f=(0, 0, 0), halo=(0, 1, 1), in_ld=192, in_coff=64, out_ld=320, out_coff=64, out_halo=(0, 1, 1))
    run_slab_case("T12 1x1x1 s(1,2,2) 256->512 odd 55x55 (i3res50 layer2.0.downsample)", K, 2, (2, 55, 55), 256, 256, 512, (1, 1, 1),
                  stride=(1, 2, 2), pad_f=(0, 0, 0))
    run_slab_case("T13 1x1x1 s(2,2,2) 64->128 (r3d layer2.0.downsample)", K, 2, (4, 28, 28), 64, 64, 128, (1, 1, 1), stride=(2, 2, 2),
                  pad_f=(0, 0, 0))
    run_slab_case("T14 1x1x1 s(2,2,2) 128->256 odd depth 5x14x14", K, 2, (5, 14, 14), 128, 128, 256, (1, 1, 1), stride=(2, 2, 2),
                  pad_f=(0, 0, 0))


def group_fuzz():
    """Seeded random shapes through the engine's own kind selection (resident / pair / stream / stream-pair, stacked
    rows, odd extents, slices, fused pool, residual) against torch."""
    import random
    from tedspad_b200 import engine
    # TEDSPAD_FUZZ_SEED / TEDSPAD_FUZZ_N: other seeds and longer runs (profiles/r2j_fuzz_extended.txt); the defaults are the
    # battery of the driver-run GPU suite
    for seed0 in [int(v) for v in os.environ.get("TEDSPAD_FUZZ_SEED", "1234").split(",")]:
        _fuzz_seed(seed0, int(os.environ.get("TEDSPAD_FUZZ_N", "48")), random, engine)


def _fuzz_seed(seed0, count, random, engine):
    rnd = random.Random(seed0)
    rnd_s = random.Random(seed0 ^ 0x5151)   # strides come from their own stream: the shapes of a seed do not depend on them
    tag = "" if seed0 == 1234 else f"s{seed0}."
    for i in range(count):
        threeD = rnd.random() < 0.35
        sp = rnd.choice([(3, 3), (3, 3), (1, 1)])
        kd = rnd.choice([1, 3]) if threeD else 1
        k = (kd, sp[0], sp[1])
        cin = rnd.choice([64, 64, 128, 192, 256])
        cout = rnd.choice([32, 64, 64, 96, 128, 160, 256, 320, 512])
        N = rnd.randint(1, 5)
        D = rnd.randint(2, 5) if threeD else 1
        H, W = rnd.randint(5, 61), rnd.randint(5, 61)
        halo = (0, 1, 1) if (not threeD and rnd.random() < 0.6) else (0, 0, 0)
        res = rnd.random() < 0.25
        pool = (not res) and (not threeD) and sp == (3, 3) and H % 2 == 0 and W % 2 == 0 and rnd.random() < 0.4
        pad = (kd // 2, sp[0] // 2, sp[1] // 2)
        # strided 1x1x1 (ResNet down-sample projections: the streaming kind with a strided TMA box)
        stride = (1, 1, 1)
        if k == (1, 1, 1) and rnd_s.random() < 0.5:
            stride = (2, 2, 2) if (threeD and rnd_s.random() < 0.5) else (1, 2, 2)
            pool = False
        # pick the kind the executors would pick for this layer
        g = torch.Generator(device="cpu").manual_seed(i)
        wtmp = torch.zeros(cout, cin, *k)
        pc = ops.PackedConv(wtmp, None, None, stride=stride, pad_front=pad, cin_pad=cin, device=DEV, n_align=32)
        ps = engine.slab3x3(pc)
        if ps is None:
            print(f"[SKIP] fuzz {i}: no slab kind for k={k} {cin}->{cout}")
            continue
        slice_out = rnd.random() < 0.3
        run_slab_case(f"Z{tag}{i} k={k} {cin}->{cout} N={N} D={D} {H}x{W} halo={halo[1]} kind={ps.kind}"
                      f"{' s=' + str(stride) if stride != (1, 1, 1) else ''}"
                      f"{' +pool' if pool else ''}{' +res' if res else ''}{' slice' if slice_out else ''}",
                      ps.kind, N, (D, H, W), cin, cin, cout, k, stride=stride, pad_f=pad, halo=halo, pool=pool, res=res, seed=100 + i,
                      out_ld=(cout + 32) if slice_out else None, out_coff=16 if slice_out else 0)


def group_streampair():
    """Streaming kind on CTA pairs: every CTA streams half of each weight block's rows."""
    K = L.SLAB_3X3_STREAM_PAIR
    run_slab_case("Q1 128->128 20x24 haloed pair", K, 2, (1, 20, 24), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("Q2 256->128 28x28 tm1 pair", K, 2, (1, 28, 28), 256, 256, 128, (1, 3, 3), halo=(0, 1, 1), tm=1)
    run_slab_case("Q3 256->128 slices pair", K, 4, (1, 32, 32), 256, 256, 128, (1, 3, 3), halo=(0, 1, 1), in_ld=320, in_coff=64,
                  out_ld=256, out_coff=128)
    run_slab_case("Q4 3x3x3 64->192 no halo pair", K, 2, (4, 28, 20), 64, 64, 192, (3, 3, 3), pad_f=(1, 1, 1))
    run_slab_case("Q5 3x3x3 96(128)->208 -> n_tile 224 pair", K, 2, (4, 14, 14), 96, 128, 224, (3, 3, 3), pad_f=(1, 1, 1),
                  out_ld=480, out_coff=192)
    run_slab_case("Q6 (1,3,3) 256->256 28x28 pair (N=256)", K, 2, (1, 28, 28), 256, 256, 256, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("Q7 128->128 112x112 x8 many tiles pair", K, 8, (1, 112, 112), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("Q8 128->128 pool fused pair", K, 4, (1, 32, 48), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1), pool=True)
    run_slab_case("Q9 384-out two N tiles -> single-CTA fallback", K, 4, (2, 7, 7), 192, 192, 384, (3, 3, 3), pad_f=(1, 1, 1))


def group_streampairperf():
    for K, nm in ((L.SLAB_3X3_STREAM, "single"), (L.SLAB_3X3_STREAM_PAIR, "pair")):
        time_slab(f"128->128 @112 x128 {nm}", K, 128, (1, 112, 112), 128, 128, (1, 3, 3))
        time_slab(f"128->128 @112 x128 +pool {nm}", K, 128, (1, 112, 112), 128, 128, (1, 3, 3), pool=True)
        time_slab(f"256->128 @112 x128 {nm}", K, 128, (1, 112, 112), 256, 128, (1, 3, 3))
        time_slab(f"256->128 @56 x128 {nm}", K, 128, (1, 56, 56), 256, 128, (1, 3, 3))
        time_slab(f"128->256 @56 x128 {nm}", K, 128, (1, 56, 56), 128, 256, (1, 3, 3))
        time_slab(f"512->256 @56 x128 {nm}", K, 128, (1, 56, 56), 512, 256, (1, 3, 3))
        time_slab(f"512->256 @28 x128 {nm}", K, 128, (1, 28, 28), 512, 256, (1, 3, 3))
        time_slab(f"i3d 2c 64->192 3x3x3 @8x56x56 x8 {nm}", K, 8, (8, 56, 56), 64, 192, (3, 3, 3), pad_f=(1, 1, 1))
        time_slab(f"i3d 3c.b1b 128->192 @8x28x28 x8 {nm}", K, 8, (8, 28, 28), 128, 192, (3, 3, 3), pad_f=(1, 1, 1))


def group_pairperf():
    for K, nm in ((L.SLAB_3X3, "single"), (L.SLAB_3X3_PAIR, "pair")):
        time_slab(f"64->64 @224 x128 {nm}", K, 128, (1, 224, 224), 64, 64, (1, 3, 3))
        time_slab(f"64->64 @224 x128 +pool {nm}", K, 128, (1, 224, 224), 64, 64, (1, 3, 3), pool=True)
        time_slab(f"128->64 @224 x128 {nm}", K, 128, (1, 224, 224), 128, 64, (1, 3, 3))
        time_slab(f"128->64 @112 x128 {nm}", K, 128, (1, 112, 112), 128, 64, (1, 3, 3))
        time_slab(f"64->128 @112 x128 {nm}", K, 128, (1, 112, 112), 64, 128, (1, 3, 3))


def group_e16perf():
    """Sixteen- against eight-warp epilogue on the 64-output CTA-pair layers (run with TEDSPAD_SLAB_E16=0/1/2)."""
    K = L.SLAB_3X3_PAIR
    for _ in range(2):
        time_slab("64->64 @224 x128 pair", K, 128, (1, 224, 224), 64, 64, (1, 3, 3))
        time_slab("64->64 @224 x128 +pool pair", K, 128, (1, 224, 224), 64, 64, (1, 3, 3), pool=True)
        time_slab("64->64 @224 x128 OutConv-only pair", K, 128, (1, 224, 224), 64, 64, (1, 3, 3), outconv=True)
        time_slab("128->64 @224 x128 pair", K, 128, (1, 224, 224), 128, 64, (1, 3, 3))


def group_stempairperf():
    for tm in (1, 2):
        for K, nm in ((L.SLAB_STEM3D, "single"), (L.SLAB_STEM3D_PAIR, "pair")):
            time_slab(f"stem3d i3d 7x7x7 s2 x32 tm{tm} {nm}", K, 32, (16, 224, 224), 4, 64, (7, 7, 7), stride=(2, 2, 2),
                      pad_f=(2, 2, 2), pad_b=(3, 3, 3), tm=tm, cin_real=3)
    for K, nm in ((L.SLAB_STEM3D, "single"), (L.SLAB_STEM3D_PAIR, "pair")):
        time_slab(f"stem3d i3res50 5x7x7 s2 x32 {nm}", K, 32, (16, 224, 224), 4, 64, (5, 7, 7), stride=(2, 2, 2), pad_f=(2, 3, 3),
                  cin_real=3)


def group_slabstream():
    K = L.SLAB_3X3_STREAM
    run_slab_case("R1 128->128 20x24 haloed", K, 2, (1, 20, 24), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("R2 256->128 28x28 tm1 1cta", K, 2, (1, 28, 28), 256, 256, 128, (1, 3, 3), halo=(0, 1, 1), tm=1, max_ctas=1)
    run_slab_case("R3 256->128 slices", K, 3, (1, 32, 32), 256, 256, 128, (1, 3, 3), halo=(0, 1, 1), in_ld=320, in_coff=64,
                  out_ld=256, out_coff=128)
    run_slab_case("R4 3x3x3 64->192 no halo", K, 2, (4, 28, 20), 64, 64, 192, (3, 3, 3), pad_f=(1, 1, 1))
    run_slab_case("R5 3x3x3 96(128)->208 padded K, slice out", K, 2, (4, 14, 14), 96, 128, 208, (3, 3, 3), pad_f=(1, 1, 1),
                  out_ld=480, out_coff=192)
    run_slab_case("R6 3x3x3 192->384 two N tiles 7x7", K, 4, (2, 7, 7), 192, 192, 384, (3, 3, 3), pad_f=(1, 1, 1))
    run_slab_case("R7 (1,3,3) D=2 512->512 no halo", K, 2, (2, 14, 14), 512, 512, 512, (1, 3, 3))
    run_slab_case("R8 128->128 112x112 x8 many tiles", K, 8, (1, 112, 112), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1))
    run_slab_case("R9 3x3x3 64->64 r3d 8x28x28", K, 2, (8, 28, 28), 64, 64, 64, (3, 3, 3), pad_f=(1, 1, 1))
    run_slab_case("R10 128->128 pool fused", K, 3, (1, 32, 48), 128, 128, 128, (1, 3, 3), halo=(0, 1, 1), pool=True)
    run_slab_case("R11 256->256 pool fused (N=256, tm=1)", K, 2, (1, 28, 28), 256, 256, 256, (1, 3, 3), halo=(0, 1, 1), pool=True)
    run_slab_case("R12 512->512 two N tiles pool fused", K, 2, (1, 28, 28), 512, 512, 512, (1, 3, 3), halo=(0, 1, 1), pool=True)


def run_up_case(name, kind, N, hw, c_skip, c_up, cout, up_hw=None, tm=0):
    """conv([skip | upsample2x(low)]) fused (slab producers interpolate) vs the unfused pair of kernels (bit-exact)
    and vs torch (F.interpolate align_corners=True + F.pad + cat + conv2d)."""
    try:
        g = torch.Generator(device="cpu").manual_seed(len(name))
        H, W = hw
        uh, uw = up_hw or (H // 2, W // 2)
        skip = torch.randn(N, c_skip, H, W, generator=g).to(DEV)
        low = torch.randn(N, c_up, uh, uw, generator=g).to(DEV)
        cin = c_skip + c_up
        w = (torch.randn(cout, cin, 1, 3, 3, generator=g) / (9 * cin) ** 0.5).to(DEV)
        b = (torch.rand(cout, generator=g) - 0.5).to(DEV)
        pc = ops.PackedConv(w, b, None, pad_front=(0, 1, 1), cin_pad=cin, device=DEV, n_align=32)
        psc = ops.PackedSlabConv(pc, kind)
        cat = ops.CLTensor(N, 1, H, W, cin, (0, 1, 1), device=DEV)
        cat.slice(0, c_skip).interior()[...] = skip.unsqueeze(2).permute(0, 2, 3, 4, 1).to(torch.bfloat16)
        cat.slice(c_skip, c_up).interior()[...] = 9.0   # must never be read by the fused path
        lowc = ops.CLTensor.from_ncdhw(low, halo=(0, 1, 1))
        y1 = ops.CLTensor(N, 1, H, W, cout, (0, 1, 1), device=DEV)
        ops.conv_slab_forward(cat.slice(0, c_skip), psc, y1, up=lowc, tm=tm)
        # unfused: materialise the up-sampled half, then the same convolution over the full concat buffer
        ops.upsample2x(lowc, cat.slice(c_skip, c_up))
        y2 = ops.CLTensor(N, 1, H, W, cout, (0, 1, 1), device=DEV)
        ops.conv_slab_forward(cat, psc, y2, tm=tm)
        torch.cuda.synchronize()
        same = bool(torch.equal(y1.buf, y2.buf))
        RESULTS.append((name + ":fused==unfused", same))
        print(f"[{'PASS' if same else 'FAIL'}] {name}: fused == unfused bit for bit")
        up = F.interpolate(bf(low), scale_factor=2, mode="bilinear", align_corners=True).to(torch.bfloat16).float()
        dy, dx = H - up.shape[2], W - up.shape[3]
        up = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        wq = pc.w[:cout, :9 * cin].float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
        ref = torch.relu(F.conv2d(torch.cat([bf(skip), up], 1), wq.contiguous(), pc.bias[:cout], padding=1)).unsqueeze(2)
        report(name, y1.to_ncdhw(), ref)
    except Exception:
        RESULTS.append((name, False))
        print(f"[FAIL] {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_slabup():
    run_up_case("U1 up4.0 shape 64|64->64 resident 32x48", L.SLAB_3X3, 2, (32, 48), 64, 64, 64)
    run_up_case("U2 resident tm1 odd 21x27 (F.pad)", L.SLAB_3X3, 2, (21, 27), 64, 64, 64, up_hw=(10, 13), tm=1)
    run_up_case("U3 up3.0 shape 128|128->128 stream 28x28", L.SLAB_3X3_STREAM, 3, (28, 28), 128, 128, 128)
    run_up_case("U4 up1.0 shape 512|512->512 stream 28x28 (2 N tiles)", L.SLAB_3X3_STREAM, 2, (28, 28), 512, 512, 512)
    run_up_case("U5 stream 256|256->256 56x56 x4", L.SLAB_3X3_STREAM, 4, (56, 56), 256, 256, 256)
    run_up_case("U6 resident 64|64->64 112x112 x8 many tiles", L.SLAB_3X3, 8, (112, 112), 64, 64, 64)


def group_slabstem():
    run_slab_case("T1 stem2d 3(8)->64 20x20 tm2", L.SLAB_STEM2D, 2, (1, 20, 20), 3, 8, 64, (1, 3, 3), tm=2, out_halo=(0, 1, 1))
    run_slab_case("T2 stem2d tm1 40x24", L.SLAB_STEM2D, 3, (1, 40, 24), 3, 8, 64, (1, 3, 3), tm=1, out_halo=(0, 1, 1))
    run_slab_case("T3 stem2d 112x112 x4", L.SLAB_STEM2D, 4, (1, 112, 112), 3, 8, 64, (1, 3, 3), out_halo=(0, 1, 1))
    run_slab_case("T4 stem3d i3d 7x7x7 s2 SAME", L.SLAB_STEM3D, 1, (8, 20, 24), 3, 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3))
    run_slab_case("T5 stem3d i3res50 5x7x7 s2 p(2,3,3)", L.SLAB_STEM3D, 2, (8, 32, 32), 3, 4, 64, (5, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 3, 3), tm=2)
    run_slab_case("T6 stem3d r3d (3,7,7) s(1,2,2)", L.SLAB_STEM3D, 2, (4, 28, 28), 3, 4, 64, (3, 7, 7), stride=(1, 2, 2),
                  pad_f=(1, 3, 3))
    run_slab_case("T7 stem3d i3d 16x64x64 x2", L.SLAB_STEM3D, 2, (16, 64, 64), 3, 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3))
    # 2-D 7x7 stride-2 pad-3 stem (kd = 1): the ResNet-18 encoder conv1 of the UNet++ anonymizer, frames as [N][1][H][W][4]
    run_slab_case("T8 stem 2-D 7x7 s2 p3 (resnet18 conv1) 32x48 x3", L.SLAB_STEM3D, 3, (1, 32, 48), 3, 4, 64, (1, 7, 7), stride=(1, 2, 2),
                  pad_f=(0, 3, 3), out_halo=(0, 1, 1))
    run_slab_case("T9 stem 2-D 7x7 s2 p3 112x112 x4 tm2 into a slice", L.SLAB_STEM3D, 4, (1, 112, 112), 3, 4, 64, (1, 7, 7), stride=(1, 2, 2),
                  pad_f=(0, 3, 3), tm=2, out_halo=(0, 1, 1), out_ld=384, out_coff=256)
    # the same stems on CTA pairs (cta_group::2, half of the weight rows per CTA); TP5 has an odd tile count -> fallback
    KP = L.SLAB_STEM3D_PAIR
    run_slab_case("TP1 stem3d pair i3d 7x7x7 s2 SAME", KP, 1, (8, 20, 24), 3, 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3))
    run_slab_case("TP2 stem3d pair i3res50 5x7x7 s2 p(2,3,3) tm2", KP, 2, (8, 32, 32), 3, 4, 64, (5, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 3, 3), tm=2)
    run_slab_case("TP3 stem3d pair r3d (3,7,7) s(1,2,2)", KP, 2, (4, 28, 28), 3, 4, 64, (3, 7, 7), stride=(1, 2, 2),
                  pad_f=(1, 3, 3))
    run_slab_case("TP4 stem3d pair i3d 16x64x64 x2", KP, 2, (16, 64, 64), 3, 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3))
    run_slab_case("TP5 stem pair 2-D 7x7 s2 p3 32x48 x3 (odd tiles: single-CTA fallback)", KP, 3, (1, 32, 48), 3, 4, 64, (1, 7, 7),
                  stride=(1, 2, 2), pad_f=(0, 3, 3), out_halo=(0, 1, 1))
    run_slab_case("TP6 stem pair 2-D 7x7 s2 p3 112x112 x4 tm2 into a slice", KP, 4, (1, 112, 112), 3, 4, 64, (1, 7, 7), stride=(1, 2, 2),
                  pad_f=(0, 3, 3), tm=2, out_halo=(0, 1, 1), out_ld=384, out_coff=256)
    # planar anonymizer output -> encoder clip (raw-reshape glue)
    try:
        g = torch.Generator(device="cpu").manual_seed(5)
        B, T, H, W = 2, 16, 12, 16
        fr = torch.rand(B * T, 3, H, W, generator=g).to(DEV).to(torch.bfloat16)
        ref = fr.float().reshape(B, T, 3, H, W).reshape(B, 3, T, H, W)
        for cpad in (4, 8):
            enc = ops.CLTensor(B, T, H, W, cpad, device=DEV)
            enc.buf.fill_(5.0)
            ops.planes_to_clip(fr, enc, T)
            report(f"planes_to_clip C={cpad}", enc.to_ncdhw()[:, :3], ref, tol_rel=0)
            okz = bool((enc.interior()[..., 3:] == 0).all())
            RESULTS.append((f"planes_to_clip C={cpad} pad", okz))
            print(f"[{'PASS' if okz else 'FAIL'}] planes_to_clip C={cpad}: pad channels zero")
    except Exception:
        RESULTS.append(("planes_to_clip", False))
        print(f"[FAIL] planes_to_clip: EXCEPTION\n{traceback.format_exc()}", flush=True)


def time_slab(name, kind, N, dhw, cin_buf, cout, k, stride=(1, 1, 1), pad_f=(0, 1, 1), pad_b=None, tm=0, iters=10, pool=False,
              cin_real=None, n_tile=0, outconv=False):
    try:
        D, H, W = dhw
        cin_real = cin_real or cin_buf
        halo = (0, 1, 1) if (kind in (L.SLAB_3X3, L.SLAB_3X3_STREAM, L.SLAB_3X3_PAIR, L.SLAB_3X3_STREAM_PAIR, L.SLAB_3X3_KX_PAIR) and D == 1) else (0, 0, 0)
        x = ops.CLTensor(N, D, H, W, cin_buf, halo, device=DEV)
        x.interior().normal_()
        wt = torch.randn(cout, cin_real, *k, device=DEV) / (cin_real * k[0] * k[1] * k[2]) ** 0.5
        pc = ops.PackedConv(wt, None, None, stride=stride, pad_front=pad_f,
                            cin_pad=cin_buf if kind in (L.SLAB_3X3, L.SLAB_3X3_STREAM, L.SLAB_3X3_PAIR, L.SLAB_3X3_STREAM_PAIR, L.SLAB_3X3_KX_PAIR) else 8, device=DEV, n_align=32)
        psc = ops.PackedSlabConv(pc, kind, n_tile=n_tile)
        od, oh, ow = pc.out_extent((D, H, W), pad_b)
        y = ops.CLTensor(N, od, oh, ow, cout, (0, 1, 1) if od == 1 else (0, 0, 0), device=DEV)
        pv = ops.CLTensor(N, 1, oh // 2, ow // 2, cout, (0, 1, 1), device=DEV) if pool else None
        oc, yo = None, y
        if outconv:   # the last UNet layer: OutConv + sigmoid + glue fused, the 64-channel tensor is never written
            clip = ops.CLTensor(N // 16, 16, oh, ow, 4, device=DEV)
            oc = (torch.randn(3, cout, device=DEV) / 8, torch.randn(3, device=DEV), None, None, clip, 16)
            yo = None
        for _ in range(3):
            ops.conv_slab_forward(x, psc, yo, pool=pv, tm=tm, outconv=oc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv_slab_forward(x, psc, yo, pool=pv, tm=tm, outconv=oc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * N * od * oh * ow * cin_real * cout * k[0] * k[1] * k[2]
        plan = psc.plan(x, y, tm=tm)
        print(f"[PERF] slab {name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)  tm={plan.tm} stages={plan.stages} "
              f"b_stages={plan.b_stages} n_tile={plan.n_tile}x{plan.num_n_tiles}", flush=True)
    except Exception:
        print(f"[FAIL] perf slab {name}: EXCEPTION\n{traceback.format_exc()}", flush=True)


def group_streamperf():
    K = L.SLAB_3X3_STREAM
    for tm in (1, 2):
        time_slab(f"128->128 @112 x128 tm{tm}", K, 128, (1, 112, 112), 128, 128, (1, 3, 3), tm=tm)
    time_slab("256->128 @112 x128", K, 128, (1, 112, 112), 256, 128, (1, 3, 3))
    time_slab("256->128 @56 x128", K, 128, (1, 56, 56), 256, 128, (1, 3, 3))
    time_slab("128->256 @56 x128", K, 128, (1, 56, 56), 128, 256, (1, 3, 3))
    time_slab("i3d 2c 64->192 3x3x3 @8x56x56 x8", K, 8, (8, 56, 56), 64, 192, (3, 3, 3), pad_f=(1, 1, 1))
    time_slab("i3d 3c.b1b 128->192 @8x28x28 x8", K, 8, (8, 28, 28), 128, 192, (3, 3, 3), pad_f=(1, 1, 1))
    time_slab("i3d 4f.b1b 192(160)->320 @4x14x14 x8", K, 8, (4, 14, 14), 192, 320, (3, 3, 3), pad_f=(1, 1, 1))
    time_conv("FLAT 128->128 @112 x128 (old feed)", 128, (112, 112), 128, 128)
    time_conv("FLAT 256->128 @112 x128 (old feed)", 128, (112, 112), 256, 128)


def group_wideperf():
    """256/512-channel UNet layers: streaming SLAB (128- and 256-wide N tiles) against the FLAT feed."""
    K = L.SLAB_3X3_STREAM
    for (cin, cout, hw) in ((128, 256, 56), (256, 256, 56), (512, 256, 56), (256, 512, 28), (512, 512, 28), (1024, 512, 28),
                            (512, 512, 14)):
        for nt in (128, 256):
            time_slab(f"{cin}->{cout} @{hw} x128 n_tile={nt}", K, 128, (1, hw, hw), cin, cout, (1, 3, 3), n_tile=nt)
        time_conv(f"FLAT {cin}->{cout} @{hw} x128", 128, (hw, hw), cin, cout)


def group_slabperf():
    K = L.SLAB_3X3
    for tm in (1, 2):
        time_slab(f"64->64 @224 x128 tm{tm}", K, 128, (1, 224, 224), 64, 64, (1, 3, 3), tm=tm)
    time_slab("64->64 @224 x128 +pool", K, 128, (1, 224, 224), 64, 64, (1, 3, 3), pool=True)
    time_slab("128->64 @224 x128", K, 128, (1, 224, 224), 128, 64, (1, 3, 3))
    time_slab("64->128 @112 x128", K, 128, (1, 112, 112), 64, 128, (1, 3, 3))
    time_slab("128->64 @112 x128", K, 128, (1, 112, 112), 128, 64, (1, 3, 3))
    for tm in (1, 2):
        time_slab(f"stem2d 3->64 @224 x128 tm{tm}", L.SLAB_STEM2D, 128, (1, 224, 224), 8, 64, (1, 3, 3), tm=tm, cin_real=3)
        time_slab(f"stem3d i3d 7x7x7 s2 x8 tm{tm}", L.SLAB_STEM3D, 8, (16, 224, 224), 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3), tm=tm, cin_real=3)
        time_slab(f"stem3d PAIR i3d 7x7x7 s2 x8 tm{tm}", L.SLAB_STEM3D_PAIR, 8, (16, 224, 224), 4, 64, (7, 7, 7), stride=(2, 2, 2),
                  pad_f=(2, 2, 2), pad_b=(3, 3, 3), tm=tm, cin_real=3)
    time_conv("FLAT 64->64 @224 x128 (old feed)", 128, (224, 224), 64, 64)


if __name__ == "__main__":
    groups = sys.argv[1:] or ["flat", "gather", "ops", "prep"]
    t0 = time.time()
    print(f"device: {torch.cuda.get_device_name(0)}  sms={L.lib().tedspad_num_sms()}", flush=True)
    for gname in groups:
        print(f"===== group {gname} =====", flush=True)
        globals()["group_" + gname]()
    torch.cuda.synchronize()
    nfail = sum(1 for _, ok in RESULTS if not ok)
    print(f"===== {len(RESULTS) - nfail} passed, {nfail} failed, {time.time() - t0:.1f}s =====", flush=True)
    sys.exit(1 if nfail else 0)
