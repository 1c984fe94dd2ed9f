"""Host-side wrappers of the C-ABI operators: channels-last activation buffers, weight packing
and one Python function per entry point of include/tedspad.h.

PyTorch is used here for device memory and streams only; all arithmetic is in libtedspad.so.
"""
import ctypes as C
import math

import torch

from . import _lib as L

BF16 = torch.bfloat16

LAUNCHES = 0        # kernels launched through this module (bench.py reports it as gpu_launches)
CONV_EVENTS = None  # when a list: one (start, end) CUDA event pair is appended per conv launch (bench.py roofline)
OP_EVENTS = None    # when a list: (start, end, name) per non-convolution launch (tests/profile_step.py)


def _count():
    global LAUNCHES
    LAUNCHES += 1


class _timed:
    """with _timed("maxpool"): ...  records a CUDA event pair on the current stream when OP_EVENTS is a list."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if OP_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if OP_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            OP_EVENTS.append((self.e0, e1, self.name))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what}: a CUDA tensor is required - this framework has no CPU path")


class CLTensor:
    """Channels-last activation buffer [N][D+2pd][H+2ph][W+2pw][ld] with a zero halo, or a channel
    slice [coff, coff+C) of one (zero-copy concatenation)."""

    def __init__(self, N, D, H, W, C_, halo=(0, 0, 0), device="cuda", dtype=BF16, ld=None, buf=None, coff=0):
        self.N, self.D, self.H, self.W, self.C = int(N), int(D), int(H), int(W), int(C_)
        self.pd, self.ph, self.pw = (int(h) for h in halo)
        self.ld = int(ld) if ld is not None else self.C
        self.coff = int(coff)
        shape = (self.N, self.D + 2 * self.pd, self.H + 2 * self.ph, self.W + 2 * self.pw, self.ld)
        if buf is None:
            has_halo = (self.pd + self.ph + self.pw) > 0
            buf = (torch.zeros if has_halo else torch.empty)(shape, device=device, dtype=dtype)
        assert tuple(buf.shape) == shape and buf.is_contiguous(), (tuple(buf.shape), shape)
        self.buf = buf

    @property
    def halo(self):
        return (self.pd, self.ph, self.pw)

    def slice(self, coff, C_):
        assert 0 <= coff and coff + C_ <= self.ld
        return CLTensor(self.N, self.D, self.H, self.W, C_, self.halo, ld=self.ld, buf=self.buf, coff=coff)

    def desc(self):
        return L.TensorDesc(self.buf.data_ptr(), self.N, self.D, self.H, self.W, self.C, self.pd, self.ph, self.pw,
                            self.ld, self.coff)

    def interior(self):
        """torch view [N, D, H, W, C] of the logical tensor (no copy)."""
        return self.buf[:, self.pd:self.pd + self.D, self.ph:self.ph + self.H, self.pw:self.pw + self.W,
                        self.coff:self.coff + self.C]

    def to_ncdhw(self):
        return self.interior().permute(0, 4, 1, 2, 3).float()

    @staticmethod
    def from_ncdhw(t, halo=(0, 0, 0), c_pad=8, ld=None):
        """Test/adapter helper: fp32 [N,C,D,H,W] (or [N,C,H,W]) torch tensor -> CLTensor (via torch copy)."""
        if t.dim() == 4:
            t = t.unsqueeze(2)
        N, C_, D, H, W = t.shape
        Cp = int(math.ceil(C_ / c_pad) * c_pad)
        out = CLTensor(N, D, H, W, Cp, halo, device=t.device, ld=ld)
        out.buf.zero_()
        out.interior()[..., :C_] = t.permute(0, 2, 3, 4, 1).to(BF16)
        return out


def n_tiling(cout, align=16):
    """(n_tile, Cout_pad) the conv kernels use for `cout` output channels (the SLAB feed wants align=32)."""
    c16 = (cout + align - 1) // align * align
    nt = (c16 + 255) // 256
    n_tile = ((c16 + nt - 1) // nt + align - 1) // align * align
    return n_tile, n_tile * nt


class PackedConv:
    """Convolution weights in the kernel's layout: bf16 [Cout_pad][K_pad] K-major with K ordered
    (kd,kh,kw,cin_pad); BatchNorm (eval) folded in fp32 before rounding; fp32 bias [Cout_pad]."""

    def __init__(self, weight, bias=None, bn=None, stride=(1, 1, 1), pad_front=(0, 0, 0), cin_pad=None,
                 device="cuda", n_align=16):
        w = weight.detach().float()
        if w.dim() == 4:  # Conv2d [Cout,Cin,kh,kw] -> 3-D with kd=1
            w = w.unsqueeze(2)
        if w.dim() == 2:  # Linear [out,in] -> 1x1x1
            w = w[:, :, None, None, None]
        cout, cin, kd, kh, kw = w.shape
        b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
        if bn is not None:
            gamma, beta, mean, var, eps = bn
            scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
            w = w * scale.view(-1, 1, 1, 1, 1)
            b = (b - mean.detach().float()) * scale + beta.detach().float()
        self.cin = cin
        self.cin_pad = int(cin_pad) if cin_pad is not None else (cin + 7) // 8 * 8
        assert self.cin_pad >= cin and self.cin_pad % 8 == 0
        self.cout = cout
        self.n_tile, self.cout_pad = n_tiling(cout, n_align)
        self.k = (kd, kh, kw)
        self.stride = tuple(int(s) for s in stride)
        self.pad_front = tuple(int(p) for p in pad_front)
        k_real = kd * kh * kw * self.cin_pad
        self.k_pad = (k_real + 63) // 64 * 64
        wp = torch.zeros(self.cout_pad, kd, kh, kw, self.cin_pad, dtype=torch.float32, device=w.device)
        wp[:cout, :, :, :, :cin] = w.permute(0, 2, 3, 4, 1)
        wk = torch.zeros(self.cout_pad, self.k_pad, dtype=torch.float32, device=w.device)
        wk[:, :k_real] = wp.reshape(self.cout_pad, k_real)
        self.w = wk.to(BF16).to(device).contiguous()
        bp = torch.zeros(self.cout_pad, dtype=torch.float32, device=w.device)
        bp[:cout] = b
        self.bias = bp.to(device).contiguous()

    @staticmethod
    def concat(pcs):
        """One GEMM for convolutions that read the same input (same kernel / stride / pads / cin_pad): their weight
        rows are stacked, each block padded to a multiple of 16 rows.  Returns the merged PackedConv; .seg_begin[i]
        is the first output column of pcs[i] (tedspad_conv y / y2 / y3)."""
        a = pcs[0]
        assert all((p.k, p.stride, p.pad_front, p.cin_pad, p.k_pad) == (a.k, a.stride, a.pad_front, a.cin_pad, a.k_pad)
                   for p in pcs) and 1 <= len(pcs) <= 3
        m = PackedConv.__new__(PackedConv)
        m.cin, m.cin_pad, m.k, m.stride, m.pad_front, m.k_pad = a.cin, a.cin_pad, a.k, a.stride, a.pad_front, a.k_pad
        rows = [(p.cout + 15) // 16 * 16 for p in pcs]
        m.seg_begin = [sum(rows[:i]) for i in range(len(pcs))]
        m.seg_cout = [p.cout for p in pcs]
        m.cout = m.seg_begin[-1] + pcs[-1].cout
        m.n_tile, m.cout_pad = n_tiling(sum(rows), 16)
        w = torch.zeros(m.cout_pad, m.k_pad, dtype=BF16, device=a.w.device)
        b = torch.zeros(m.cout_pad, dtype=torch.float32, device=a.bias.device)
        for p, r0 in zip(pcs, m.seg_begin):
            w[r0:r0 + p.cout] = p.w[:p.cout]
            b[r0:r0 + p.cout] = p.bias[:p.cout]
        m.w, m.bias = w.contiguous(), b.contiguous()
        return m

    def out_extent(self, in_extent, pad_back=None):
        """Output (D,H,W) for an input (D,H,W); pad_back defaults to pad_front (symmetric)."""
        pb = self.pad_front if pad_back is None else pad_back
        return tuple((i + pf + b - k) // s + 1
                     for i, k, s, pf, b in zip(in_extent, self.k, self.stride, self.pad_front, pb))


def conv_forward(x, pc, y, res=None, act=L.ACT_RELU, feed=L.FEED_AUTO, y_fp32=False, max_ctas=0, n_tile=0):
    """y = act(conv(x, pc) + bias (+ res)).  x, y, res are CLTensor views.  With a PackedConv.concat() merge, y is the
    list of destination views (one per merged convolution)."""
    assert x.C == pc.cin_pad, f"input view has {x.C} channels, weights were packed for {pc.cin_pad}"
    d = L.ConvDesc()
    if isinstance(y, (list, tuple)):
        ys = list(y)
        y = ys[0]
        assert len(ys) == len(pc.seg_begin) and res is None and all(t.C == c for t, c in zip(ys, pc.seg_cout))
        if len(ys) > 1:
            d.y2, d.y2_begin = ys[1].desc(), pc.seg_begin[1]
        if len(ys) > 2:
            d.y3, d.y3_begin = ys[2].desc(), pc.seg_begin[2]
    d.x, d.y = x.desc(), y.desc()
    d.w, d.bias = pc.w.data_ptr(), pc.bias.data_ptr()
    if res is not None:
        assert (res.N, res.D, res.H, res.W, res.halo) == (y.N, y.D, y.H, y.W, y.halo) and res.C == y.C
        d.res, d.res_ld, d.res_coff = res.buf.data_ptr(), res.ld, res.coff
    d.Cout, d.Cout_pad, d.K_pad = pc.cout, pc.cout_pad, pc.k_pad
    d.kd, d.kh, d.kw = pc.k
    d.sd, d.sh, d.sw = pc.stride
    d.pd, d.ph, d.pw = pc.pad_front
    d.act, d.y_fp32, d.feed, d.n_tile, d.max_ctas = act, int(y_fp32), feed, n_tile or pc.n_tile, max_ctas
    _count()
    if CONV_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().tedspad_conv_forward(C.byref(d), _stream()), "tedspad_conv_forward")
        e1.record()
        CONV_EVENTS.append((e0, e1, (x.N, x.D, x.H, x.W, x.C, pc.cout, pc.k, pc.stride, y.D, y.H, y.W, pc.cin)))
        return y
    L.check(L.lib().tedspad_conv_forward(C.byref(d), _stream()), "tedspad_conv_forward")
    return y


class PackedSlabConv:
    """Weights of a SLAB-feed convolution (include/tedspad.h: tedspad_conv_slab): the standard PackedConv
    layout re-packed on the device into the shared-memory image the kernel keeps resident."""

    def __init__(self, pc, kind, n_tile=0):
        self.pc, self.kind = pc, int(kind)
        # STREAM kind: UMMA N per tile (<= 256).  Measured on B200 (tests/gpu_diag.py wideperf): for 256/512 outputs
        # one 256-wide tile beats two 128-wide ones; 128 outputs get two 8-column halves per weight block.
        self.n_tile = int(n_tile) or pc.n_tile
        self.cout, self.cout_pad = pc.cout, pc.cout_pad
        self.bias = pc.bias
        self.cin_pad = 4 if kind in (L.SLAB_STEM3D, L.SLAB_STEM3D_PAIR) else pc.cin_pad
        self.fallback = None
        if kind in (L.SLAB_3X3_STREAM, L.SLAB_3X3_STREAM_PAIR):
            # weights stream from the standard packed layout: nothing to re-pack
            if pc.cout_pad % 32 or pc.cout_pad > 2048 or pc.cout % 8 or self.n_tile % 32 or pc.cout_pad % self.n_tile:
                raise ValueError(f"slab stream feed needs Cout_pad % 32 == 0 (<= 2048), got {pc.cout_pad} / n_tile {pc.n_tile}")
            self.image, self.image_bytes = pc.w, pc.w.numel() * 2
            if kind == L.SLAB_3X3_STREAM_PAIR:
                self.fallback = PackedSlabConv(pc, L.SLAB_3X3_STREAM, n_tile)
            return
        if pc.cout_pad > 256 or pc.cout % 8:
            raise ValueError(f"slab feed needs a single N tile (Cout_pad={pc.cout_pad}) with Cout % 8 == 0")
        # CTA-pair kind: layers it cannot tile (W <= 8, odd tile count) run through the single-CTA kind
        if kind == L.SLAB_3X3_PAIR:
            self.fallback = PackedSlabConv(pc, L.SLAB_3X3)
        if kind == L.SLAB_STEM3D_PAIR:
            self.fallback = PackedSlabConv(pc, L.SLAB_STEM3D)
        if kind == L.SLAB_3X3_KX_PAIR:
            # whatever the KX kind cannot run (fused pool / OutConv / residual, odd tile counts) takes the plain kinds
            self.fallback = PackedSlabConv(pc, L.SLAB_3X3_PAIR if pc.cout_pad in (64, 128) else L.SLAB_3X3)
        nbytes = C.c_int64(0)
        args = (self.kind, None, pc.cout_pad, pc.k_pad, pc.cin_pad, *pc.k, pc.pad_front[2])
        L.check(L.lib().tedspad_conv_slab_pack(*args, None, C.byref(nbytes), None), "tedspad_conv_slab_pack(size)")
        self.image_bytes = int(nbytes.value)
        self.image = None
        if pc.w.is_cuda:
            self.image = torch.empty(self.image_bytes // 2, dtype=BF16, device=pc.w.device)
            args = (self.kind, pc.w.data_ptr(), pc.cout_pad, pc.k_pad, pc.cin_pad, *pc.k, pc.pad_front[2])
            _count()
            L.check(L.lib().tedspad_conv_slab_pack(*args, self.image.data_ptr(), C.byref(nbytes), _stream()),
                    "tedspad_conv_slab_pack")

    def desc(self, x, y, act=L.ACT_RELU, pool=None, outconv=None, tm=0, max_ctas=0, up=None, stack_rows=0, res=None,
             s2d_clip=None):
        """outconv = (w fp32 [3,Cout], b fp32 [3], planes bf16 [N,3,H,W] | None, frames fp32 [N,3,H,W] | None
        [, clip CLTensor [B,T,H,W,>=3], T]): with `clip` the sigmoid images go straight into the encoder input
        through the raw-reshape glue"""
        pc = self.pc
        d = L.ConvSlabDesc()
        d.x = x.desc()
        if y is not None:
            d.y = y.desc()
        else:  # fused OutConv only: extents of the (never written) convolution output
            d.y = L.TensorDesc(None, x.N, x.D, x.H, x.W, pc.cout, 0, 0, 0, pc.cout, 0)
        d.w_image = self.image.data_ptr() if self.image is not None else None
        d.bias = self.bias.data_ptr()
        if pool is not None:
            d.pool = pool.desc()
        if up is not None:
            d.up = up.desc()
        if res is not None:
            assert (res.N, res.D, res.H, res.W, res.halo, res.C) == (y.N, y.D, y.H, y.W, y.halo, y.C)
            d.res, d.res_ld, d.res_coff = res.buf.data_ptr(), res.ld, res.coff
        if outconv is not None:
            w, b, planes, frames = outconv[:4]
            d.oc_w, d.oc_b = w.data_ptr(), b.data_ptr()
            d.oc_planes = planes.data_ptr() if planes is not None else None
            d.oc_frames = frames.data_ptr() if frames is not None else None
            if len(outconv) > 4:
                d.oc_clip, d.oc_T = outconv[4].desc(), int(outconv[5])
        if s2d_clip is not None:   # KX kind: (clip CLTensor [B,T,2H,2W,>=3], T) - the space-to-depth head's glue, fused
            d.oc_clip, d.oc_T = s2d_clip[0].desc(), int(s2d_clip[1])
        d.kind, d.Cout, d.Cout_pad = self.kind, pc.cout, pc.cout_pad
        d.kd, d.kh, d.kw = pc.k
        d.sd, d.sh, d.sw = pc.stride
        d.pd, d.ph, d.pw = pc.pad_front
        d.act, d.tm, d.max_ctas, d.stack_rows = act, tm, max_ctas, stack_rows
        d.n_tile, d.K_pad = (self.n_tile if self.kind in (L.SLAB_3X3_STREAM, L.SLAB_3X3_STREAM_PAIR) else 0), pc.k_pad
        return d

    def resolve(self, x, tm=0, up=None, stack_rows=0, y=None, pool=None):
        """The PackedSlabConv that runs this input: the CTA-pair kind needs 16x16 tiles (W > 8) in an even number
        (per-image row tiles: the pair layers run at 224 / 112 where those are exact)."""
        if self.kind == L.SLAB_3X3_PAIR:
            if up is not None or tm == 1 or x.W <= 8:
                return self.fallback
            # the plan knows the tile count (stacked rows or per-image row tiles): an odd one cannot be split over pairs
            d = self.desc(x, None, tm=tm, stack_rows=stack_rows, pool=pool)   # (a fused pool constrains the row stacking)
            if L.lib().tedspad_conv_slab_plan(C.byref(d), C.byref(L.SlabPlan())) != 0:
                return self.fallback
        if self.kind == L.SLAB_3X3_KX_PAIR:
            if up is not None or pool is not None or tm != 0 or stack_rows > 0:
                return self.fallback.resolve(x, tm, up, stack_rows, y, pool)
            d = self.desc(x, None)
            if L.lib().tedspad_conv_slab_plan(C.byref(d), C.byref(L.SlabPlan())) != 0:
                return self.fallback.resolve(x, tm, up, stack_rows, y, pool)
        if self.kind == L.SLAB_STEM3D_PAIR and y is not None:
            # an odd tile count cannot be split over CTA pairs (tiles = 16 rows x 8*tm columns of one output plane)
            if (y.N * y.D * (-(-y.H // 16)) * (-(-y.W // (8 * (tm or 1))))) % 2:
                return self.fallback
        if self.kind == L.SLAB_3X3_STREAM_PAIR:
            # the plan knows (stacked rows, tile shape): an odd tile count cannot be split over CTA pairs; 3-D tensors
            # (few tiles per launch) were measured 0-5 % slower on pairs
            if up is not None or x.D > 1:
                return self.fallback
            d = self.desc(x, None, tm=tm, stack_rows=stack_rows)   # (the plan needs the output extents only)
            if L.lib().tedspad_conv_slab_plan(C.byref(d), C.byref(L.SlabPlan())) != 0:
                return self.fallback
        return self

    def plan(self, x, y, **kw):
        """The kernel's tiling / descriptor plan (host-only call; used by the CPU simulator tests)."""
        plan = L.SlabPlan()
        d = self.resolve(x, kw.get("tm", 0), kw.get("up"), kw.get("stack_rows", 0), y, kw.get("pool")).desc(x, y, **kw)
        L.check(L.lib().tedspad_conv_slab_plan(C.byref(d), C.byref(plan)), "tedspad_conv_slab_plan")
        return plan


def conv_slab_forward(x, psc, y, act=L.ACT_RELU, pool=None, outconv=None, tm=0, max_ctas=0, up=None, stack_rows=0, res=None,
                      s2d_clip=None):
    """y = act(conv(x) + bias) through the SLAB feed; optional fused MaxPool2d(2) -> pool, OutConv 1x1 + sigmoid
    -> planar images (y may then be None), and fused Up.forward input: conv([x | upsample2x(up)]).  s2d_clip = (clip,
    T): the KX kind scatters its 12-channel space-to-depth output straight into the encoder clip (y may be None); the
    caller checks with slab_runs_kx() that the KX kind really runs this input."""
    if psc.kind == L.SLAB_3X3_KX_PAIR and (outconv is not None or res is not None):
        psc = psc.fallback
    psc = psc.resolve(x, tm, up, stack_rows, y, pool)
    if s2d_clip is not None and psc.kind != L.SLAB_3X3_KX_PAIR:
        raise RuntimeError("conv_slab_forward: the fused space-to-depth clip glue needs the KX kind (check slab_runs_kx first)")
    d = psc.desc(x, y, act=act, pool=pool, outconv=outconv, tm=tm, max_ctas=max_ctas, up=up, stack_rows=stack_rows, res=res,
                 s2d_clip=s2d_clip)
    _count()
    if CONV_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().tedspad_conv_slab_forward(C.byref(d), _stream()), "tedspad_conv_slab_forward")
        e1.record()
        pc = psc.pc
        od, oh, ow = (y.D, y.H, y.W) if y is not None else (x.D, x.H, x.W)
        CONV_EVENTS.append((e0, e1, (x.N, x.D, x.H, x.W, x.C, pc.cout, pc.k, pc.stride, od, oh, ow, pc.cin)))
        return y
    L.check(L.lib().tedspad_conv_slab_forward(C.byref(d), _stream()), "tedspad_conv_slab_forward")
    return y


def slab_runs_kx(x, psc):
    """True when `psc` is a KX-kind convolution that runs input x as such (even tile count, no fallback)."""
    return psc is not None and psc.kind == L.SLAB_3X3_KX_PAIR and psc.resolve(x) is psc


def planes_to_clip(planes, y, T):
    """planar bf16 [B*T,3,H,W] anonymizer output -> encoder input view [B,T,H,W,4|8] (raw-reshape glue)."""
    _require_cuda(planes, "planes_to_clip")
    assert planes.dtype == BF16 and planes.is_contiguous()
    yd = y.desc()
    _count()
    with _timed("planes_to_clip"):
        L.check(L.lib().tedspad_planes_to_clip(planes.data_ptr(), C.byref(yd), int(T), _stream()), "tedspad_planes_to_clip")
    return y


def maxpool(x, y, k, s, pad_front=(0, 0, 0), zero_pad=False):
    xd, yd = x.desc(), y.desc()
    _count()
    with _timed("maxpool"):
        L.check(L.lib().tedspad_maxpool(C.byref(xd), C.byref(yd), *k, *s, *pad_front, int(zero_pad), _stream()),
                "tedspad_maxpool")
    return y


def upsample2x(x, y):
    xd, yd = x.desc(), y.desc()
    _count()
    with _timed("upsample2x"):
        L.check(L.lib().tedspad_upsample2x(C.byref(xd), C.byref(yd), _stream()), "tedspad_upsample2x")
    return y


def upsample2x_nearest(x, y):
    """y = nearest x2 of x, written into the channel slice y (smp DecoderBlock.forward's F.interpolate + torch.cat)."""
    xd, yd = x.desc(), y.desc()
    _count()
    with _timed("upsample2x_nearest"):
        L.check(L.lib().tedspad_upsample2x_nearest(C.byref(xd), C.byref(yd), _stream()), "tedspad_upsample2x_nearest")
    return y


def frames_to_clip(x, y, T, frames_out=None, s2d=False):
    """channels-last anonymizer frames [B*T,1,H,W,>=3] (s2d: space-to-depth [B*T,1,H/2,W/2,>=12]) -> encoder clip view
    y [B,T,H,W,4|8] (raw-reshape glue)."""
    xd, yd = x.desc(), y.desc()
    fo = frames_out.data_ptr() if frames_out is not None else None
    _count()
    with _timed("frames_to_clip"):
        L.check(L.lib().tedspad_frames_to_clip(C.byref(xd), C.byref(yd), int(T), int(s2d), fo, _stream()),
                "tedspad_frames_to_clip")
    return y


def outconv_sigmoid(x, w, b, y, T, frames_out=None):
    """w: fp32 [3, C] cuda, b: fp32 [3] cuda; y: encoder-input CLTensor [B,T,H,W,>=3]."""
    xd, yd = x.desc(), y.desc()
    fo = frames_out.data_ptr() if frames_out is not None else None
    _count()
    with _timed("outconv"):
        L.check(L.lib().tedspad_outconv_sigmoid(C.byref(xd), w.data_ptr(), b.data_ptr(), C.byref(yd), int(T), fo,
                                                _stream()), "tedspad_outconv_sigmoid")
    return y


def avgpool_features(x, kd=0):
    od = x.D - (kd if kd > 0 else x.D) + 1
    out = torch.empty((x.N, od, x.C), device=x.buf.device, dtype=torch.float32)
    xd = x.desc()
    _count()
    with _timed("avgpool"):
        L.check(L.lib().tedspad_avgpool_features(C.byref(xd), int(kd), out.data_ptr(), _stream()),
                "tedspad_avgpool_features")
    return out


def l2_normalize_rows(x, eps=1e-12):
    """In-place x / max(||x||_2, eps) per row of a contiguous fp32 cuda [rows, cols] tensor (F.normalize(p=2, dim=1))."""
    _require_cuda(x, "l2_normalize_rows")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    _count()
    L.check(L.lib().tedspad_l2_normalize_rows(x.data_ptr(), int(x.shape[0]), int(x.shape[1]), float(eps), _stream()),
            "tedspad_l2_normalize_rows")
    return x


def linear(x_f32, pc, name_bufs, act=L.ACT_NONE):
    """fp32 cuda [B, in] -> fp32 [B, out] through the convolution kernel as a 1x1x1 layer (nn.Linear, optionally with a
    folded BatchNorm1d and ReLU): `pc` = PackedConv(weight [out,in], bias, bn); name_bufs = (_Buffers, name)."""
    bufs, name = name_bufs
    B, cin = x_f32.shape
    fin = bufs.get(name + ".in", B, 1, 1, 1, pc.cin_pad)
    nchw_to_cl(x_f32.reshape(B, cin, 1, 1, 1), fin)
    out = bufs.get(name + ".out", B, 1, 1, 1, pc.cout, dtype=torch.float32)
    conv_forward(fin, pc, out, act=act, y_fp32=True)
    return out.buf.reshape(B, pc.cout)


def preprocess(frames_u8, desc_i32, crop_hw, y, resample=L.RESAMPLE_AA_FLOAT, frames_f32=None):
    """frames_u8: cuda uint8 [F,Hs,Ws,3]; desc_i32: cuda int32 [n_out,4] = (src_frame, top, left, hflip)."""
    _require_cuda(frames_u8, "preprocess")
    frames_u8 = frames_u8.contiguous()
    F_, Hs, Ws, ch = frames_u8.shape
    assert ch == 3 and frames_u8.dtype == torch.uint8
    assert desc_i32.dtype == torch.int32 and desc_i32.is_contiguous() and desc_i32.shape[1] == 4
    yd = y.desc()
    fo = frames_f32.data_ptr() if frames_f32 is not None else None
    _count()
    with _timed("preprocess"):
        L.check(L.lib().tedspad_preprocess(frames_u8.data_ptr(), F_, Hs, Ws, desc_i32.data_ptr(), desc_i32.shape[0],
                                           int(crop_hw[0]), int(crop_hw[1]), C.byref(yd), int(resample), fo, _stream()),
                "tedspad_preprocess")
    return y


def nchw_to_cl(x_f32, y):
    """fp32 [N,C,H,W] / [N,C,D,H,W] cuda tensor -> channels-last bf16 view y (extra channels zeroed)."""
    _require_cuda(x_f32, "nchw_to_cl")
    x_f32 = x_f32.contiguous().float()
    yd = y.desc()
    _count()
    L.check(L.lib().tedspad_nchw_to_cl(x_f32.data_ptr(), int(x_f32.shape[1]), C.byref(yd), _stream()),
            "tedspad_nchw_to_cl")
    return y
