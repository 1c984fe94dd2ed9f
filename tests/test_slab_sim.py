"""CPU tests of the SLAB feed: the plan the C-ABI library builds (TMA boxes, UMMA descriptor fields, MMA table)
and the weight image, replayed with the hardware addressing rules in tests/_slabsim.py, must reproduce
F.conv3d.  This pins tiling, tap offsets, swizzles, overlapped descriptors and K padding without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))

import _slabsim as S  # noqa: E402
from tedspad_b200 import _lib as L, ops  # noqa: E402


def _bf(t):
    return t.to(torch.bfloat16).float()


def _run(kind, x_cl, x_ref, w, stride, pad_f, pad_b, y_shape, tm, tiles=None, cin_pad=8, estr=None):
    """x_cl: CLTensor view (cpu) holding x_ref's data; returns max abs error over the simulated tiles."""
    cout = w.shape[0]
    b = torch.linspace(-0.5, 0.5, cout)
    pc = ops.PackedConv(w, b, None, stride=stride, pad_front=pad_f, cin_pad=cin_pad, device="cpu", n_align=32)
    psc = ops.PackedSlabConv(pc, kind)
    y = ops.CLTensor(*y_shape, cout, device="cpu")
    plan = psc.plan(x_cl, y, tm=tm)
    if kind == L.SLAB_3X3_STREAM:
        image = S.bf16_bits(pc.w)
    else:
        image = S.pack_image(kind, S.bf16_bits(pc.w), pc.cout_pad, pc.k_pad, pc.cin_pad, pc.k, pad_f[2])
        assert image.size * 2 == psc.image_bytes == plan.w_bytes * (2 if plan.pair else 1)
    wq = pc.w[:cout, :pc.k[0] * pc.k[1] * pc.k[2] * pc.cin_pad].float().reshape(cout, *pc.k, pc.cin_pad)
    wq = wq[..., :w.shape[1]].permute(0, 4, 1, 2, 3).contiguous()
    pad6 = (pad_f[2], pad_b[2], pad_f[1], pad_b[1], pad_f[0], pad_b[0])
    ref = F.conv3d(F.pad(_bf(x_ref), pad6), wq, pc.bias[:cout], stride=stride)  # [N,Cout,OD,OH,OW]
    assert tuple(ref.shape[2:]) == tuple(y_shape[1:]), (ref.shape, y_shape)
    tiles = list(range(plan.total_tiles)) if tiles is None else [t % plan.total_tiles for t in tiles]
    res = S.simulate_tiles(plan, S.bf16_bits(x_cl.buf), image, pc.bias.numpy(), tiles, k_pad=pc.k_pad, estr=estr)
    worst, seen = 0.0, 0
    OH, OW = y_shape[2], y_shape[3]
    for tile, (n, tz, oy, ox, acc, n0) in res.items():
        ok = (oy >= 0) & (oy < OH) & (ox < OW)
        nc = min(plan.n_tile, cout - n0)   # real channels of this N tile
        want = ref[torch.from_numpy(n[ok]), n0:n0 + nc, tz, torch.from_numpy(oy[ok]), torch.from_numpy(ox[ok])].numpy()
        worst = max(worst, float(np.abs(acc[ok][:, :nc] - want).max()))
        seen += int(ok.sum()) if n0 == 0 else 0
    return worst, seen, plan


@pytest.mark.parametrize("cin,cout,tm,ld,coff", [(64, 64, 2, 64, 0), (128, 64, 0, 192, 64), (64, 128, 1, 64, 0)])
def test_slab_3x3_plan_reproduces_conv(cin, cout, tm, ld, coff):
    g = torch.Generator().manual_seed(cin + cout)
    N, H, W = 2, 20, 24
    x = torch.randn(N, cin, 1, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 3, 3, generator=g) / (9 * cin) ** 0.5
    buf = ops.CLTensor(N, 1, H, W, ld, (0, 1, 1), device="cpu")
    if ld != cin:
        buf.interior()[...] = 7.0  # junk outside the view must never be read
    xv = buf.slice(coff, cin)
    xv.interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    err, seen, plan = _run(L.SLAB_3X3, xv, x, w, (1, 1, 1), (0, 1, 1), (0, 1, 1), (N, 1, H, W), tm, cin_pad=cin)
    assert plan.k_stages == cin // 64 and plan.n_mma == 36 and plan.swizzle128 == 1
    assert plan.tm == (tm if tm else 1)  # 128->64 weights leave room for three stages only at tm=1
    assert seen == N * H * W and err < 2e-5, (err, seen)


@pytest.mark.parametrize("name,dhw,cin,cout,kd,halo,tm", [
    ("unet 128->128 haloed", (1, 20, 24), 128, 128, 1, (0, 1, 1), 0),
    ("unet 256->128 tm1", (1, 18, 16), 256, 128, 1, (0, 1, 1), 1),
    ("i3d Conv3d_2c 64->192 3x3x3 no halo", (4, 20, 12), 64, 192, 3, (0, 0, 0), 0),
    ("i3d 5c.b1b 192->384 (two N tiles)", (2, 7, 7), 192, 384, 3, (0, 0, 0), 0),
    ("i3d 4b.b1b 128(96)->208 padded", (3, 14, 14), 128, 208, 3, (0, 0, 0), 0),
    ("1x1x1 bottleneck conv1 256->64 odd", (2, 13, 11), 256, 64, -1, (0, 0, 0), 0),
    ("(3,1,1) temporal conv1 128->64", (4, 9, 12), 128, 64, -3, (0, 0, 0), 0),
    ("2-D 1x1 128->128 haloed (stacked rows)", (1, 20, 24), 128, 128, -1, (0, 1, 1), 0),
])
def test_slab_stream_plan_reproduces_conv(name, dhw, cin, cout, kd, halo, tm):
    g = torch.Generator().manual_seed(cin + cout + kd)
    N = 2
    D, H, W = dhw
    x = torch.randn(N, cin, D, H, W, generator=g)
    ks = 3
    if kd < 0:   # negative kd: |kd| temporal taps with a 1x1 spatial window (one tap per K stage)
        kd, ks = -kd, 1
    w = torch.randn(cout, cin, kd, ks, ks, generator=g) / (ks * ks * kd * cin) ** 0.5
    xc = ops.CLTensor(N, D, H, W, cin, halo, device="cpu")
    xc.buf.zero_()
    xc.interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    pad = (kd // 2, ks // 2, ks // 2)
    tiles = None if D * H * W < 2000 else list(range(0, 10 ** 6, 7))[:24]
    err, seen, plan = _run(L.SLAB_3X3_STREAM, xc, x, w, (1, 1, 1), pad, pad, (N, D, H, W), tm, tiles=tiles, cin_pad=cin)
    assert plan.b_stream == 1 and plan.k_stages == kd * (cin // 64) and plan.b_stages >= 3
    assert plan.num_n_tiles == (2 if cout > 256 else 1)
    if tiles is None:
        assert seen == N * D * H * W
    assert err < 3e-5, (name, err)


@pytest.mark.parametrize("name,dhw,cin,cout,stride", [
    ("i3res50 layer2.0.downsample 256->512 s(1,2,2) odd 13x11", (2, 13, 11), 256, 512, (1, 2, 2)),
    ("r3d layer2.0.downsample 64->128 s(2,2,2) odd depth", (5, 14, 10), 64, 128, (2, 2, 2)),
    ("resnet18 layer2.0.downsample 2-D 64->128 s2, 23x21", (1, 23, 21), 64, 128, (1, 2, 2)),
])
def test_slab_stream_strided_1x1_plan_reproduces_conv(name, dhw, cin, cout, stride):
    """Strided 1x1x1 (the ResNet down-sample projections): the plan's tile origins are in INPUT pixels and the TMA box
    walks the input with the convolution's stride (conv_slab.cu forward: elementStrides = (1, sw, sh, 1, 1))."""
    g = torch.Generator().manual_seed(cin + cout + sum(stride))
    N = 2
    D, H, W = dhw
    x = torch.randn(N, cin, D, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 1, 1, generator=g) / cin ** 0.5
    xc = ops.CLTensor(N, D, H, W, cin, (0, 0, 0), device="cpu")
    xc.interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    oshape = (N, (D - 1) // stride[0] + 1, (H - 1) // stride[1] + 1, (W - 1) // stride[2] + 1)
    err, seen, plan = _run(L.SLAB_3X3_STREAM, xc, x, w, stride, (0, 0, 0), (0, 0, 0), oshape, 0, cin_pad=cin,
                           estr=(1, stride[2], stride[1], 1, 1))
    assert plan.b_stream == 1 and plan.k_stages == cin // 64 and plan.stack_hp == 0
    assert plan.x_step == 8 * plan.tm * stride[2] and plan.y_step == 16 * stride[1] and plan.z_step == stride[0]
    assert seen == oshape[0] * oshape[1] * oshape[2] * oshape[3] and err < 3e-5, (name, err, seen)


def test_slab_stem2d_plan_reproduces_conv():
    g = torch.Generator().manual_seed(3)
    N, H, W = 2, 20, 20
    x = torch.rand(N, 3, 1, H, W, generator=g)
    w = torch.randn(64, 3, 1, 3, 3, generator=g) / 27 ** 0.5
    xc = ops.CLTensor(N, 1, H, W, 8, device="cpu")
    xc.buf.zero_()
    xc.interior()[..., :3] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    for tm in (1, 2):
        err, seen, plan = _run(L.SLAB_STEM2D, xc, x, w, (1, 1, 1), (0, 1, 1), (0, 1, 1), (N, 1, H, W), tm)
        assert plan.a_layout == 0 and plan.a_lbo == 16 and plan.n_mma == 6 and plan.k_stages == 1
        assert seen == N * H * W and err < 2e-5, (tm, err)


@pytest.mark.parametrize("name,kd,sd,pad_f,pad_b,dhw", [
    ("i3d Conv3d_1a_7x7 TF-SAME", 7, 2, (2, 2, 2), (3, 3, 3), (8, 20, 24)),
    ("I3Res50 conv1", 5, 2, (2, 3, 3), (2, 3, 3), (8, 20, 24)),
    ("r3d_18 stem", 3, 1, (1, 3, 3), (1, 3, 3), (4, 20, 24)),
])
def test_slab_stem3d_plan_reproduces_conv(name, kd, sd, pad_f, pad_b, dhw):
    g = torch.Generator().manual_seed(kd)
    N = 1
    D, H, W = dhw
    x = torch.rand(N, 3, D, H, W, generator=g)
    w = torch.randn(64, 3, kd, 7, 7, generator=g) / (147 * kd) ** 0.5
    xc = ops.CLTensor(N, D, H, W, 4, device="cpu")
    xc.buf.zero_()
    xc.interior()[..., :3] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    od = (D + pad_f[0] + pad_b[0] - kd) // sd + 1
    oh = (H + pad_f[1] + pad_b[1] - 7) // 2 + 1
    ow = (W + pad_f[2] + pad_b[2] - 7) // 2 + 1
    for tm in (1, 2):
        err, seen, plan = _run(L.SLAB_STEM3D, xc, x, w, (sd, 2, 2), pad_f, pad_b, (N, od, oh, ow), tm)
        assert plan.k_stages == kd and plan.n_mma == 14
        assert seen == N * od * oh * ow and err < 2e-5, (name, tm, err)
    # the same stem on CTA pairs: two per-CTA weight images of 32 rows (when the tile count is even)
    psc = ops.PackedSlabConv(ops.PackedConv(w, None, None, stride=(sd, 2, 2), pad_front=pad_f, cin_pad=8, device="cpu", n_align=32),
                             L.SLAB_STEM3D_PAIR)
    y = ops.CLTensor(N, od, oh, ow, 64, device="cpu")
    if psc.resolve(xc, y=y) is psc:
        err, seen, plan = _run(L.SLAB_STEM3D_PAIR, xc, x, w, (sd, 2, 2), pad_f, pad_b, (N, od, oh, ow), 0)
        assert plan.pair == 1 and plan.b_lbo == 32 * 16
        assert seen == N * od * oh * ow and err < 2e-5, (name, "pair", err)


def test_slab_plan_rejects_what_it_cannot_run():
    pc = ops.PackedConv(torch.zeros(64, 64, 1, 3, 3), None, None, pad_front=(0, 1, 1), device="cpu")
    psc = ops.PackedSlabConv(pc, L.SLAB_3X3)
    x_c32 = ops.CLTensor(1, 1, 16, 16, 32, device="cpu")
    y = ops.CLTensor(1, 1, 16, 16, 64, device="cpu")
    with pytest.raises(RuntimeError, match="multiple of 64"):
        psc.plan(x_c32, y)
    big = ops.PackedConv(torch.zeros(128, 128, 1, 3, 3), None, None, pad_front=(0, 1, 1), device="cpu")
    x = ops.CLTensor(1, 1, 16, 16, 128, (0, 1, 1), device="cpu")
    with pytest.raises(RuntimeError, match="do not fit"):
        ops.PackedSlabConv(big, L.SLAB_3X3).plan(x, ops.CLTensor(1, 1, 16, 16, 128, device="cpu"))


@pytest.mark.parametrize("cin,cout,N,H,W", [(64, 64, 2, 20, 24), (128, 16, 4, 16, 30), (64, 40, 2, 19, 21)])
def test_slab_kx_plan_reproduces_conv(cin, cout, N, H, W):
    """KX kind: plan (16-pixel-wide slab, SBO 1024, three filter-row groups, N = 3 x Cout_pad), numpy packer and the
    epilogue's neighbour sums reproduce the convolution on every pixel (ragged rows / columns, stacked rows)."""
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(N, cin, 1, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.linspace(-0.5, 0.5, cout)
    xc = ops.CLTensor(N, 1, H, W, cin, (0, 1, 1), device="cpu")
    xc.interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    pc = ops.PackedConv(w, b, None, pad_front=(0, 1, 1), cin_pad=cin, device="cpu", n_align=32)
    psc = ops.PackedSlabConv(pc, L.SLAB_3X3_KX_PAIR)
    y = ops.CLTensor(N, 1, H, W, cout, (0, 1, 1), device="cpu")
    plan = psc.plan(xc, y)
    assert plan.pair == 1 and plan.tm == 1 and plan.n_tile == 3 * pc.cout_pad and plan.n_grp == 3 and plan.total_tiles % 2 == 0
    image = S.pack_image(L.SLAB_3X3_KX_PAIR, S.bf16_bits(pc.w), pc.cout_pad, pc.k_pad, pc.cin_pad, pc.k, 1)
    assert image.size * 2 == psc.image_bytes == 2 * plan.w_bytes
    wq = pc.w[:cout, :9 * cin].float().reshape(cout, 1, 3, 3, cin).permute(0, 4, 1, 2, 3).contiguous()
    ref = F.conv3d(F.pad(_bf(x), (1, 1, 1, 1, 0, 0)), wq, pc.bias[:cout])[:, :, 0]   # [N,Cout,H,W]
    res = S.simulate_tiles_kx(plan, S.bf16_bits(xc.buf), image, pc.bias.numpy(), range(plan.total_tiles), pc.cout_pad)
    worst, seen = 0.0, 0
    for tile, (n, oy, ox, out) in res.items():
        ok = (oy >= 0) & (oy < H) & (ox < W)
        want = ref[torch.from_numpy(n[ok]), :, torch.from_numpy(oy[ok]), torch.from_numpy(ox[ok])].numpy()
        worst = max(worst, float(np.abs(out[ok][:, :cout] - want).max()))
        seen += int(ok.sum())
    assert seen == N * H * W and worst < 3e-5, (worst, seen)


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 64), (64, 128)])
def test_slab_3x3_pair_plan_reproduces_conv(cin, cout):
    """CTA-pair kind: per-CTA weight images of Cout_pad / 2 rows at the same table offsets (stacked rows, 20 x 24)."""
    g = torch.Generator().manual_seed(3 * cin + cout)
    N, H, W = 2, 20, 24
    x = torch.randn(N, cin, 1, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 3, 3, generator=g) / (9 * cin) ** 0.5
    xc = ops.CLTensor(N, 1, H, W, cin, (0, 1, 1), device="cpu")
    xc.interior()[...] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    err, seen, plan = _run(L.SLAB_3X3_PAIR, xc, x, w, (1, 1, 1), (0, 1, 1), (0, 1, 1), (N, 1, H, W), 0, cin_pad=cin)
    assert plan.pair == 1 and plan.tm == 2 and plan.total_tiles % 2 == 0
    assert seen == N * H * W and err < 2e-5, (err, seen)


def test_staged_store_tile_is_the_swizzle_128b_image_of_the_box():
    """The staged epilogue store (conv_slab.cu: SlabKParams::tmY): a warp's 32 pixels (4 tile rows x 8 columns, lane =
    row * 8 + column) x 64 channels, written chunk-swizzled by the lanes, must be exactly what a SWIZZLE_128B tensor store
    of box {64, 8, 4} reads back as [row][column][channel]; and the eight lanes of a quarter warp must hit eight distinct
    16-byte bank groups (conflict-free shared-memory stores)."""
    rs = np.random.RandomState(5)
    q = rs.randint(0, 65536, (32, 64)).astype(np.uint16)
    out = S.tma_store_box(S.stage_rows(q))
    assert out.shape == (4, 8, 64)
    assert np.array_equal(out.reshape(32, 64), q)
    for k in range(8):
        for quarter in range(4):
            lanes = np.arange(quarter * 8, quarter * 8 + 8)
            groups = (k ^ (lanes & 7)) % 8          # 16-byte column of the 128-byte row = bank group
            assert len(set(groups.tolist())) == 8

