"""Network executors: the anonymizer UNet and the three video encoders expressed as sequences of
C-ABI operator calls over channels-last bf16 buffers.

Each executor is built from a reference-format `state_dict` (the boundary nn.Modules in
aux_code/ own the parameters and hand them over), folds BatchNorm into the packed weights once,
and caches its activation buffers per input shape.  Concatenations (unet_parts.py:67, i3d.py:149)
are never materialised: producers write straight into channel slices of the consumer's input.
"""
import os

import torch

from . import _lib as L
from . import ops
from .ops import CLTensor, PackedConv

# SLAB feed switches (csrc/conv_slab.cu): on by default; "0" routes the layer classes back through the FLAT /
# GATHER feeds of csrc/conv_igemm.cu (kept for A/B measurements and as the general-shape path).
USE_SLAB = os.environ.get("TEDSPAD_SLAB", "1") != "0"
USE_PAIR = os.environ.get("TEDSPAD_PAIR", "1") != "0"     # cta_group::2 for the 64-output-channel 3x3 layers
# 1x1x1 / (3,1,1) stride-1 convolutions (ResNet bottleneck conv1 / conv3, Inception pool branch, Conv3d_2b) through the
# streaming SLAB kind (one tap per K stage) instead of the general kernel, whose 4-warp 16-column epilogue runs a
# 128 x 256 tile in ~27k clocks (ncu: I3Res50.layer1 conv3 + residual at 37 TFLOP/s)
SLAB_1X1 = USE_SLAB and os.environ.get("TEDSPAD_SLAB_1X1", "1") != "0"
PAD_SMALL_3X3 = USE_SLAB and os.environ.get("TEDSPAD_PAD_SMALL_3X3", "1") != "0"
MERGE_1X1 = os.environ.get("TEDSPAD_MERGE_1X1", "1") != "0"  # Inception b0/b1a/b2a (same input) as one GEMM
USE_STREAM_PAIR = USE_PAIR and os.environ.get("TEDSPAD_STREAM_PAIR", "1") != "0"
# Inception blocks: the pool branch (reads the block input only) and the 5x5-slot branch on side streams next to the
# heads GEMM / the 3x3 branch.  Inside a captured CUDA graph these are parallel branches (no host cost); the Mixed_4x /
# 5x launches are 15-50 us kernels of 100-400 tiles that leave SMs idle in their tails.
I3D_BRANCH_STREAMS = os.environ.get("TEDSPAD_I3D_BRANCH_STREAMS", "0") != "0"
# Inception heads (b0 / b1a / b2a: the three 1x1x1 convolutions that read the block input, i3d.py:144-148) as ONE
# single-destination convolution through the slab kernel: the block buffer is laid out physically as
# [b1b | b2b | b3b | b0 | b1a (padded) | b2a (padded)], so the heads' outputs are one contiguous channel range and the
# next block's weights are packed with their input channels permuted to that order (I3DExecutor._mixed_packed).
I3D_HEADS_SLAB = SLAB_1X1 and PAD_SMALL_3X3 and os.environ.get("TEDSPAD_I3D_HEADS_SLAB", "1") != "0"
USE_SLAB_STEM3D = USE_SLAB and os.environ.get("TEDSPAD_SLAB_STEM3D", "1") != "0"
UNETPP_FUSE_GLUE = os.environ.get("TEDSPAD_UNETPP_FUSE_GLUE", "1") != "0"   # raw-reshape glue in the head's KX epilogue
SLAB_STRIDED_1X1 = os.environ.get("TEDSPAD_SLAB_STRIDED_1X1", "1") != "0"   # down-sample projections through the slab kernel
USE_KX = USE_PAIR and os.environ.get("TEDSPAD_KX", "1") != "0"      # 3x3 layers with few outputs through the KX kind
KX_COUT_PADS = tuple(int(v) for v in os.environ.get("TEDSPAD_KX_COUT_PADS", "32").split(",") if v)
USE_STEM_PAIR = USE_PAIR and os.environ.get("TEDSPAD_STEM_PAIR", "1") != "0"   # cta_group::2 for the 64-output 7x7 stems
STEM_PAIR_MIN_KD = int(os.environ.get("TEDSPAD_STEM_PAIR_MIN_KD", "1"))
SLAB_WEIGHT_LIMIT = 150 * 1024   # bytes of resident weights that still leave room for three slab stages
ENC_IN_CHANNELS = 4 if USE_SLAB_STEM3D else 8   # channel padding of the encoder input clip
# Up.forward inside the convolution: decoder levels (1 = deepest, up1) whose up-sampled input is interpolated by the
# conv's slab producers instead of being materialised by tedspad_upsample2x.  Bit-identical to the two-kernel path
# (tests/gpu_diag.py slabup) but MEASURED SLOWER on B200 at every level (905 clips/s unfused; 890 / 881 / 845 / 751
# with levels 1 / 1-2 / 1-3 / 1-4 fused): the interpolation competes with the epilogue warps for issue slots while the
# stand-alone kernel streams its coalesced stores independently of the tensor pipe.  Off by default; kept as an option for narrower batches.
FUSE_UPSAMPLE_LEVELS = tuple(int(v) for v in os.environ.get("TEDSPAD_FUSE_UPSAMPLE", "").split(",") if v) if USE_SLAB else ()


# CUDA graphs: a forward over fixed buffers is a fixed list of ~100 kernel launches whose host side (ctypes descriptors,
# plans, tensor-map encodes, cudaLaunchKernelEx) costs 2-4 ms - more than the GPU needs for one clip.  The second call
# with the same shapes captures the launches (programmatic-dependent-launch edges included) and every later call is
# one cudaGraphLaunch.  TEDSPAD_GRAPHS=0 switches back to eager launches.
USE_GRAPHS = os.environ.get("TEDSPAD_GRAPHS", "1") != "0"


class GraphCache:
    """key -> captured CUDA graph of a launch sequence over fixed device buffers.

    run(key, gen, fn): `fn()` enqueues the launches on the current stream and returns its output tensor(s), which must
    be the SAME device memory on every call (static buffers); `gen` is a hashable that changes whenever any buffer fn
    touches may have moved (_Buffers.generation).  1st call with a key: eager.  2nd: capture + replay.  Later: replay."""

    def __init__(self):
        self.entries = {}

    def run(self, key, gen, fn):
        if not USE_GRAPHS or ops.CONV_EVENTS is not None or ops.OP_EVENTS is not None or torch.cuda.is_current_stream_capturing():
            return fn()
        ent = self.entries.get(key)
        if ent is None or ent[0] != gen:
            out = fn()                       # eager: allocates buffers, builds plans, opts kernels in to large smem
            self.entries[key] = (gen, None, None, 1)
            return out
        _, graph, out, seen = ent
        if graph is None:
            graph = torch.cuda.CUDAGraph()
            launches = ops.LAUNCHES
            try:
                # thread_local: CUDA calls of OTHER host threads (a pin-memory worker, NVML sampling) must not abort it
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = fn()
            except Exception:
                ops.LAUNCHES = launches
                self.entries[key] = (("eager-only", gen), None, None, 0)    # never matches: stays on eager launches
                raise
            self.entries[key] = (gen, graph, out, ops.LAUNCHES - launches)
            ops.LAUNCHES = launches
            seen = self.entries[key][3]
        graph.replay()
        ops.LAUNCHES += seen                 # kernels launched by the replay (bench.py's gpu_launches)
        return out


def slab3x3(pc, max_stream_cout=2048):
    """Best SLAB variant for a stride-1 same-padded (1|3) x (3x3 | 1x1) convolution with Cin % 64 == 0, or None:
    resident weights when they fit in shared memory (the 64-channel DoubleConv layers), streamed weight blocks
    otherwise (128-channel DoubleConv layers, Conv3d_2c_3x3, Inception 3x3x3 branches, ResNet (1,3,3)/3x3x3)."""
    kd, sp = pc.k[0], pc.k[1:]
    if USE_SLAB and SLAB_1X1 and SLAB_STRIDED_1X1 and pc.k == (1, 1, 1) and pc.stride != (1, 1, 1) and max(pc.stride) <= 2 and \
            pc.pad_front == (0, 0, 0) and pc.cin_pad % 64 == 0 and pc.cout % 8 == 0 and pc.cout_pad % 32 == 0 and \
            pc.n_tile % 32 == 0 and pc.cout_pad <= max_stream_cout and pc.k_pad == pc.cin_pad:
        # strided 1x1x1 (the ResNets' down-sample projections): the streaming kind with a strided TMA box
        return ops.PackedSlabConv(pc, L.SLAB_3X3_STREAM)
    if not USE_SLAB or kd not in (1, 3) or sp not in ((3, 3), (1, 1)) or pc.stride != (1, 1, 1) or \
            pc.pad_front != (kd // 2, sp[0] // 2, sp[1] // 2):
        return None
    if sp == (1, 1) and not SLAB_1X1:
        return None
    if pc.cin_pad % 64 or pc.cout % 8 or pc.cout_pad % 32 or pc.k_pad != kd * sp[0] * sp[1] * pc.cin_pad:
        return None
    if USE_KX and kd == 1 and sp == (3, 3) and pc.cout_pad in KX_COUT_PADS and 9 * pc.cin_pad * pc.cout_pad * 2 <= SLAB_WEIGHT_LIMIT:
        # few outputs: the MMA is bound by its A-operand reads; the KX kind fetches every slab pixel 3 instead of 9 times
        # (its fallback chain covers the fused-epilogue layers and the shapes it cannot tile).  Measured on B200
        # (profiles/r2d_kx_kind.txt): the 128 -> 12(16) head of the UNet++ anonymizer 0.72 -> 0.42 ms per 32 clips; at 64
        # outputs (N = 192) it is NOT faster than the CTA-pair kind - an N = 192 MMA costs what N = 256 does
        return ops.PackedSlabConv(pc, L.SLAB_3X3_KX_PAIR)
    if kd == 1 and sp == (3, 3) and pc.cout_pad <= 256 and 9 * pc.cin_pad * pc.cout_pad * 2 <= SLAB_WEIGHT_LIMIT:
        # N = 64 is shared-memory-read bound on one SM (67 % of the tensor peak): CTA pairs split the weight rows.
        # At N = 128 the halved weight image makes room for 16x16 tiles (measured +7 %).
        return ops.PackedSlabConv(pc, L.SLAB_3X3_PAIR if (USE_PAIR and pc.cout_pad in (64, 128)) else L.SLAB_3X3)
    if pc.cout_pad <= max_stream_cout and pc.n_tile % 32 == 0:
        # 2-D layers with one N tile: CTA pairs stream half a weight block each (measured +11..17 % at N = 128 / 256 on
        # 112^2 / 56^2 maps, neutral at 28^2; the 3-D Inception branches with their few tiles per launch lose 0-5 %)
        if USE_STREAM_PAIR and kd == 1 and pc.cout_pad <= 256 and pc.n_tile == pc.cout_pad:
            return ops.PackedSlabConv(pc, L.SLAB_3X3_STREAM_PAIR)
        return ops.PackedSlabConv(pc, L.SLAB_3X3_STREAM)
    return None


def conv_auto(x, pc, y, res=None, act=L.ACT_RELU):
    """Convolution through the SLAB feed when the layer has one (pc.slab; bf16 residual added in its epilogue), else
    FLAT/GATHER."""
    ps = getattr(pc, "slab", None)
    if ps is not None and act in (L.ACT_RELU, L.ACT_NONE):
        return ops.conv_slab_forward(x, ps, y, act=act, res=res)
    return ops.conv_forward(x, pc, y, res=res, act=act)


def stem3d(pc):
    """PackedSlabConv for a (kd,7,7) stride-(sd,2,2) Cin=3 stem fed from a 4-channel clip, else None."""
    if not USE_SLAB_STEM3D or pc.k[1:] != (7, 7) or pc.stride[1:] != (2, 2) or pc.cin > 4 or pc.cout_pad > 256:
        return None
    # N = 64 stems are bound by shared-memory operand reads on one SM like every 64-output layer: CTA pairs.  Measured per
    # 32 clips (profiles/r2b_stem3d_pair.txt): 7x7x7 0.56 -> 0.43 ms, 5x7x7 0.41 -> 0.32, (3,7,7) 0.59 -> 0.50-0.57, the
    # 2-D 7x7 stem of the UNet++ encoder (one K stage per tile) 0.29-0.31 -> 0.27-0.28
    pair = USE_STEM_PAIR and pc.cout_pad == 64 and pc.k[0] >= STEM_PAIR_MIN_KD
    return ops.PackedSlabConv(pc, L.SLAB_STEM3D_PAIR if pair else L.SLAB_STEM3D)


def stem_conv(x, pc, ps, y, act=L.ACT_RELU):
    """First convolution of an encoder: SLAB stem when the clip is stored with 4 channels, GATHER feed for 8."""
    if x.C == 4:
        if ps is None:
            raise RuntimeError("a 4-channel encoder clip needs the SLAB stem (TEDSPAD_SLAB_STEM3D=1)")
        return ops.conv_slab_forward(x, ps, y, act=act)
    return ops.conv_forward(x, pc, y, act=act)


def _bn(sd, prefix, eps):
    return (sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"], sd[prefix + ".running_var"], eps)


def same_pad(size, k, s):
    """TF-"SAME" padding of i3d.py:82-86: total pad, split front = total // 2."""
    total = max(k - s, 0) if size % s == 0 else max(k - size % s, 0)
    return total // 2, total - total // 2


class _Buffers:
    """Named activation buffers of one executor.  ONE allocation per name, sized for the largest batch seen so far: a
    smaller batch (the tail of a video, a flush at a resolution change) runs on the N-prefix view of the same memory,
    a larger batch or a different geometry replaces the allocation.  The footprint is therefore bounded by the largest
    batch (~0.8 GB per 16x224x224 clip for the UNet, 26 GB at 32 clips) however many distinct batch sizes a dataset's
    video tails produce."""

    def __init__(self, device):
        self.device = device
        self.pool = {}      # name -> (spec, N allocated, CLTensor | torch.Tensor)
        self.views = {}     # (name, N) -> prefix view
        self.generation = 0  # bumped by every (re)allocation: captured CUDA graphs over these buffers are then stale

    def get(self, name, N, D, H, W, C, halo=(0, 0, 0), dtype=ops.BF16, zero=False):
        """zero=True: cleared once at allocation (channel-padded buffers whose pad channels are never written)."""
        v = self.views.get((name, N))
        spec = (D, H, W, C, tuple(halo), dtype)
        if v is not None and v[0] == spec:
            return v[1]
        ent = self.pool.get(name)
        if ent is None or ent[0] != spec or ent[1] < N:
            for k in [k for k in self.views if k[0] == name]:
                del self.views[k]
            self.pool.pop(name, None)   # (freed before the replacement is allocated)
            ent = None
            self.generation += 1
            t = CLTensor(N, D, H, W, C, halo, device=self.device, dtype=dtype)
            if zero:
                t.buf.zero_()
            ent = (spec, N, t)
            self.pool[name] = ent
        full = ent[2]
        t = full if ent[1] == N else CLTensor(N, D, H, W, C, halo, ld=full.ld, buf=full.buf[:N], coff=full.coff)
        self.views[(name, N)] = (spec, t)
        return t

    def raw(self, name, shape, dtype):
        """Plain (non channels-last) scratch tensor, e.g. the planar anonymizer output."""
        key = ("raw", name)
        ent = self.pool.get(key)
        if ent is None or ent[0] != (tuple(shape), dtype):
            self.generation += 1
            ent = ((tuple(shape), dtype), 0, torch.empty(tuple(shape), device=self.device, dtype=dtype))
            self.pool[key] = ent
        return ent[2]

    def find(self, name, N=None):
        """The buffer last handed out under `name` (its N-prefix view when N is given), or None: lets the parity tests
        read intermediate activations after a run."""
        ent = self.pool.get(name)
        if ent is None:
            return None
        if N is None or N == ent[1]:
            return ent[2]
        v = self.views.get((name, N))
        return v[1] if v is not None else None

    def nbytes(self):
        return sum((e[2].buf if isinstance(e[2], CLTensor) else e[2]).numel() *
                   (e[2].buf if isinstance(e[2], CLTensor) else e[2]).element_size() for e in self.pool.values())


# ======================================================================================== UNet
class UNetExecutor:
    """aux_code/models/unet_model.py:26-37 as 18 tcgen05 convolutions + up-samples (+ pools / OutConv when they
    are not fused).  Every convolution runs through the SLAB feed: the Cin=3 stem, resident weights for the
    64-channel layers (MaxPool2d and OutConv+sigmoid fused into their epilogues), streamed weights for the rest;
    TEDSPAD_SLAB=0 switches back to the FLAT TMA / GATHER feeds over the same zero-haloed buffers."""

    HALO = (0, 1, 1)
    LEVELS = ["inc.double_conv"] + [f"down{i}.maxpool_conv.1.double_conv" for i in range(1, 5)]

    def __init__(self, sd, device):
        self.device = device
        self.bufs = _Buffers(device)
        self.graphs = GraphCache()
        self.convs = {}
        self.slabs = {}

        def dc(prefix, cin_pad0):
            a = PackedConv(sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"], _bn(sd, f"{prefix}.1", 1e-5),
                           pad_front=(0, 1, 1), cin_pad=cin_pad0, device=device)
            b = PackedConv(sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"], _bn(sd, f"{prefix}.4", 1e-5),
                           pad_front=(0, 1, 1), device=device)
            self.convs[prefix] = (a, b)
            sa = ops.PackedSlabConv(a, L.SLAB_STEM2D) if (USE_SLAB and cin_pad0 == 8) else slab3x3(a)
            self.slabs[prefix] = (sa, slab3x3(b))

        dc("inc.double_conv", 8)
        for i in range(1, 5):
            dc(f"down{i}.maxpool_conv.1.double_conv", None)
        for i in range(1, 5):
            dc(f"up{i}.conv.double_conv", None)
        self.out_w = sd["outc.conv.weight"].detach().float().reshape(3, -1).contiguous().to(device)
        self.out_b = sd["outc.conv.bias"].detach().float().contiguous().to(device)

    def input_buffer(self, n_frames, H, W):
        """The [n_frames,1,H,W,8] bf16 buffer preprocessing / nchw_to_cl writes the frames into."""
        return self.bufs.get("x0", n_frames, 1, H, W, 8)

    @staticmethod
    def _conv(x, pc, ps, y, pool=None):
        """DoubleConv half: conv+BN+ReLU (unet_parts.py:15-22) and, when `pool` is given, the MaxPool2d(2) of the
        next Down block (unet_parts.py:33) - fused into the SLAB epilogue, a separate kernel otherwise."""
        if ps is not None:
            return ops.conv_slab_forward(x, ps, y, pool=pool)
        ops.conv_forward(x, pc, y, feed=L.FEED_GATHER if pc.cin_pad == 8 else L.FEED_AUTO)
        if pool is not None:
            ops.maxpool(y, pool, (1, 2, 2), (1, 2, 2))
        return y

    def run(self, x0, enc_in, T=16, frames_out=None):
        """x0: input_buffer() filled with frames; enc_in: encoder input [B,T,H,W,4|8] (glue target)."""
        N, H, W = x0.N, x0.H, x0.W
        g, hl = self.bufs.get, self.HALO
        sizes = [(H, W)]
        for _ in range(4):
            sizes.append((sizes[-1][0] // 2, sizes[-1][1] // 2))
        ch = [64, 128, 256, 512, 512]
        # encoder path; skip tensors x1..x4 live in the first half of the concat buffers
        cats = [g(f"cat{i}", N, 1, sizes[i][0], sizes[i][1], 2 * ch[i], hl) for i in range(4)]
        cur_in = x0
        for i, prefix in enumerate(self.LEVELS):
            (a, b), (sa, sb) = self.convs[prefix], self.slabs[prefix]
            h, w = sizes[i]
            t = g(f"t{i}", N, 1, h, w, ch[i], hl)
            self._conv(cur_in, a, sa, t)
            cur = cats[i].slice(0, ch[i]) if i < 4 else g("x5", N, 1, h, w, ch[4], hl)
            nxt = g(f"p{i + 1}", N, 1, sizes[i + 1][0], sizes[i + 1][1], ch[i], hl) if i < 4 else None
            self._conv(t, b, sb, cur, pool=nxt)
            cur_in = nxt
        # decoder path
        out_ch = [256, 128, 64, 64]
        for j in range(4):
            lvl = 3 - j
            cat = cats[lvl]
            prefix = f"up{j + 1}.conv.double_conv"
            (a, b), (sa, sb) = self.convs[prefix], self.slabs[prefix]
            h, w = sizes[lvl]
            t = g(f"u{j}a", N, 1, h, w, a.cout, hl)
            if sa is not None and (j + 1) in FUSE_UPSAMPLE_LEVELS and cur.C % 64 == 0:
                # Up.forward (unet_parts.py:57-67): the up-sampled half of the concatenation is interpolated inside
                # the convolution's slab producers and never materialised
                ops.conv_slab_forward(cat.slice(0, ch[lvl]), sa, t, up=cur)
            else:
                ops.upsample2x(cur, cat.slice(ch[lvl], cur.C))
                self._conv(cat, a, sa, t)
            if j == 3 and sb is not None and enc_in.W % 8 == 0 and enc_in.C in (4, 8):
                # OutConv 1x1 + sigmoid (unet_parts.py:71-77, unet_model.py:36-37) in the last epilogue: the 64-channel
                # tensor is never written; planar images then go through the raw-reshape glue (dali_extraction.py:173)
                # ... and the sigmoid images go straight into the encoder clip through the raw-reshape glue
                # (dali_extraction.py:173): no planar intermediate, no separate glue kernel
                ops.conv_slab_forward(t, sb, None, outconv=(self.out_w, self.out_b, None, frames_out, enc_in, T))
                return enc_in
            cur = g(f"u{j}", N, 1, h, w, out_ch[j], hl)
            self._conv(t, b, sb, cur)
        ops.outconv_sigmoid(cur, self.out_w, self.out_b, enc_in, T, frames_out)
        return enc_in


# ===================================================================================== UNet++
# smp 0.3.3 UnetPlusPlus(resnet18, encoder_depth=4, decoder_channels=(256,128,64,32)) as model_loaders.py:19-30 builds it
RESNET18_LAYERS = [(64, 1), (128, 2), (256, 2), (512, 2)]   # (planes, stride of block 0); layer4 is parameters only
UNETPP_BLOCKS = [  # (name, up-sampled input channels, skip channels, output channels); see oracle/models.py
    ("x_0_0", 256, 128, 256), ("x_0_1", 256, 128, 128), ("x_1_1", 128, 64, 64),
    ("x_0_2", 128, 192, 64), ("x_1_2", 64, 128, 64), ("x_2_2", 64, 64, 64), ("x_0_3", 64, 0, 32),
]


def conv_after_nearest_up(w):
    """3x3 weights [O,I,3,3] that act on nearest_x2(low) -> 3x3 weights [4*O, I, 3, 3] that act on `low` itself and
    produce the result in space-to-depth form (output channel (2a + b)*O + o of low-res pixel (i, j) = channel o of
    pixel (2i + a, 2j + b)): row 2i + a + ky - 1 of the up-sampled image is low row i + floor((a + ky - 1) / 2), so the
    taps that fall on the same low-res pixel are summed (in fp32, before the bf16 rounding of the packed weights).
    Zero padding is preserved: high-res row -1 / 2H is low-res row -1 / H."""
    O, I = w.shape[:2]
    out = w.new_zeros((2, 2, O, I, 3, 3), dtype=torch.float32)
    wf = w.detach().float()
    for a in range(2):
        for ky in range(3):
            dy = (a + ky - 1) // 2
            for b in range(2):
                for kx in range(3):
                    out[a, b, :, :, dy + 1, (b + kx - 1) // 2 + 1] += wf[:, :, ky, kx]
    return out.reshape(4 * O, I, 3, 3)


def conv_space_to_depth(w):
    """3x3 weights [O,I,3,3] at full resolution -> [4*O, 4*I, 3, 3] on space-to-depth tensors (channel (2a + b)*C + c of
    pixel (i, j) = channel c of pixel (2i + a, 2j + b)): the same convolution, on a quarter of the pixels with four times
    the channels - every K element is a real channel (a 32-channel tensor stored for the 64-channel K blocks of the
    slab feed wastes half of the operand traffic that bounds a 224^2 layer) and N is 128 instead of 32."""
    O, I = w.shape[:2]
    out = w.new_zeros((2, 2, O, 2, 2, I, 3, 3), dtype=torch.float32)
    wf = w.detach().float()
    for a in range(2):
        for ky in range(3):
            t = a + ky - 1
            dy, a2 = t // 2, t % 2
            for b in range(2):
                for kx in range(3):
                    u = b + kx - 1
                    out[a, b, :, a2, u % 2, :, dy + 1, u // 2 + 1] += wf[:, :, ky, kx]
    return out.reshape(4 * O, 4 * I, 3, 3)


class UNetPPExecutor:
    """arch='unet++' anonymizer (aux_code/model_loaders.py:18-30): ResNet-18 encoder to layer3 + the nested UNet++
    decoder + 3x3 head, per frame.

    Dense skips without copies: per resolution ONE buffer holds every tensor that is ever concatenated there, ordered
    so that each decoder block's input is a contiguous channel range (the block's conv1 weights are packed with their
    input channels permuted to that physical order), and every producer writes straight into its slice:

        /8   P8  = [ up(f/16) 256 | f/8 128 ]                                   x_0_0 reads [0,384)
        /4   P4  = [ up(x_0_0) 256 | x_1_1 64 | f/4 64 | up(f/8) 128 ]          x_1_1 reads [320,512), x_0_1 reads [0,384)
        /2   P2  = [ up(x_0_1) 128 | x_1_2 64 | x_2_2 64 | f/2 64 | E 64 ]      x_2_2 reads [256,384) with E = up(f/4),
                                                                                 x_1_2 reads [192,384) with E = up(x_1_1)
                                                                                 (E is rewritten once x_2_2 is done),
                                                                                 x_0_2 reads [0,320)


    The full-resolution tail (x_0_3: nearest x2 of x_0_2, two 3x3 convolutions with 32 channels, then the 3x3 head) is
    computed at HALF resolution in space-to-depth form (conv_after_nearest_up / conv_space_to_depth): 64 -> 128,
    128 -> 128, 128 -> 12 channels on H/2 x W/2 pixels.  Same arithmetic, but no up-sampled tensor, K made of real
    channels only and N = 128: measured 4.6 + 0.7 ms -> 2.3 ms per 32-clip step (profiles/r2_*).  The glue kernel reads
    the 12 channels back as the 2x2 pixels they are.

    The only materialised glue is the nearest x2 up-sampling of the six inner blocks (tedspad_upsample2x_nearest)."""

    HALO = (0, 1, 1)

    def __init__(self, sd, device):
        self.device = device
        self.bufs = _Buffers(device)
        self.graphs = GraphCache()

        def mk(wk, bnk, stride=(1, 1, 1), pad=(0, 1, 1), cin_pad=None, perm=None):
            w = sd[wk]
            if perm is not None:
                w = w[:, perm]
            pc = PackedConv(w, None, _bn(sd, bnk, 1e-5), stride=stride, pad_front=pad, cin_pad=cin_pad, device=device, n_align=32)
            pc.slab = slab3x3(pc)
            return pc

        self.stem = PackedConv(sd["encoder.conv1.weight"], None, _bn(sd, "encoder.bn1", 1e-5), stride=(1, 2, 2),
                               pad_front=(0, 3, 3), cin_pad=8, device=device, n_align=32)
        self.stem_slab = stem3d(self.stem)
        self.enc = []
        inpl = 64
        for li, (planes, stride) in enumerate(RESNET18_LAYERS[:3], 1):
            for b in range(2):
                p = f"encoder.layer{li}.{b}"
                st = stride if b == 0 else 1
                has_ds = b == 0 and (st != 1 or inpl != planes)
                self.enc.append({
                    "c1": mk(f"{p}.conv1.weight", f"{p}.bn1", (1, st, st)),
                    "c2": mk(f"{p}.conv2.weight", f"{p}.bn2"),
                    "ds": mk(f"{p}.downsample.0.weight", f"{p}.downsample.1", (1, st, st), (0, 0, 0)) if has_ds else None,
                    "name": p, "li": li, "b": b, "planes": planes, "stride": st})
                inpl = planes
        r = lambda a, b: list(range(a, b))  # noqa: E731
        # physical input-channel order of conv1 per block (None = smp's own [up-sampled | skips] order)
        perms = {"x_1_1": r(128, 192) + r(0, 128),              # [f/4 | up(f/8)]
                 "x_2_2": r(64, 128) + r(0, 64),                # [f/2 | up(f/4)]
                 "x_1_2": r(64, 192) + r(0, 64)}                # [x_2_2 | f/2 | up(x_1_1)]
        self.dec = {}
        for name, cin, cskip, cout in UNETPP_BLOCKS[:-1]:
            p = f"decoder.blocks.{name}"
            self.dec[name] = (mk(f"{p}.conv1.0.weight", f"{p}.conv1.1", perm=perms.get(name)),
                              mk(f"{p}.conv2.0.weight", f"{p}.conv2.1"))
        # ---- the full-resolution tail in space-to-depth form (see the class comment)
        def mk4(w4, bnk):   # BatchNorm parameters replicated over the four pixel phases
            g, b, m, v, eps = _bn(sd, bnk, 1e-5)
            pc = PackedConv(w4, None, (g.repeat(4), b.repeat(4), m.repeat(4), v.repeat(4), eps), pad_front=(0, 1, 1),
                            device=device, n_align=32)
            pc.slab = slab3x3(pc)
            return pc
        p = "decoder.blocks.x_0_3"
        self.tail1 = mk4(conv_after_nearest_up(sd[f"{p}.conv1.0.weight"]), f"{p}.conv1.1")       # 64 -> 4 x 32
        self.tail2 = mk4(conv_space_to_depth(sd[f"{p}.conv2.0.weight"]), f"{p}.conv2.1")         # 4 x 32 -> 4 x 32
        # 3x3 head 32 -> 3 with bias, no activation: 4 x 32 -> 4 x 3 (+ 4 zero rows: 16 outputs, Cout % 8 == 0)
        hw = conv_space_to_depth(sd["segmentation_head.0.weight"])
        w16 = torch.zeros((16,) + tuple(hw.shape[1:]), dtype=torch.float32, device=hw.device)
        b16 = torch.zeros(16, dtype=torch.float32, device=hw.device)
        w16[:12], b16[:12] = hw, sd["segmentation_head.0.bias"].detach().float().repeat(4)
        self.head = PackedConv(w16, b16, None, pad_front=(0, 1, 1), device=device, n_align=32)
        self.head.slab = slab3x3(self.head)

    def input_buffer(self, n_frames, H, W):
        """The [n_frames,1,H,W,4|8] bf16 buffer preprocessing / nchw_to_cl writes the frames into."""
        return self.bufs.get("x0", n_frames, 1, H, W, 4 if self.stem_slab is not None else 8)

    def run(self, x0, enc_in, T=16, frames_out=None):
        """x0: input_buffer() filled with frames; enc_in: encoder input [B,T,H,W,4|8] (raw-reshape glue target)."""
        N, H, W = x0.N, x0.H, x0.W
        if H % 16 or W % 16:
            raise RuntimeError(f"Wrong input shape height={H}, width={W}. Expected image height and width divisible by 16.")
        g, hl = self.bufs.get, self.HALO
        sz = {k: (H // k, W // k) for k in (1, 2, 4, 8, 16)}
        P2 = g("P2", N, 1, *sz[2], 384, hl)
        P4 = g("P4", N, 1, *sz[4], 512, hl)
        P8 = g("P8", N, 1, *sz[8], 384, hl)
        f2, f4, f8 = P2.slice(256, 64), P4.slice(320, 64), P8.slice(256, 128)
        f16 = g("f16", N, 1, *sz[16], 256, hl)
        # ---- encoder: conv1 + bn1 + relu, maxpool 3x3 s2 p1 (post-ReLU input: zero padding == -inf padding), layer1..3
        stem_conv(x0, self.stem, self.stem_slab, f2)
        x = ops.maxpool(f2, g("mp", N, 1, *sz[4], 64, hl), (1, 3, 3), (1, 2, 2), (0, 1, 1), zero_pad=True)
        outs = {1: f4, 2: f8, 3: f16}
        for blk in self.enc:
            n, li, planes = blk["name"], blk["li"], blk["planes"]
            oh, ow = sz[4 << (li - 1)]
            t = conv_auto(x, blk["c1"], g(n + ".t", N, 1, oh, ow, planes, hl))
            res = x if blk["ds"] is None else conv_auto(x, blk["ds"], g(n + ".ds", N, 1, oh, ow, planes, hl), act=L.ACT_NONE)
            y = outs[li] if blk["b"] == 1 else g(n + ".y", N, 1, oh, ow, planes, hl)
            x = conv_auto(t, blk["c2"], y, res=res)      # bn2 + identity, then ReLU (torchvision BasicBlock.forward)
        # ---- decoder
        def block(name, x_in, out):
            a, b = self.dec[name]
            t = conv_auto(x_in, a, g(name + ".t", N, 1, x_in.H, x_in.W, a.cout, hl))
            return conv_auto(t, b, out)

        up = ops.upsample2x_nearest
        up(f16, P8.slice(0, 256))
        x00 = block("x_0_0", P8, g("x_0_0", N, 1, *sz[8], 256, hl))
        up(f8, P4.slice(384, 128))
        x11 = block("x_1_1", P4.slice(320, 192), P4.slice(256, 64))
        up(f4, P2.slice(320, 64))
        block("x_2_2", P2.slice(256, 128), P2.slice(192, 64))
        up(x00, P4.slice(0, 256))
        x01 = block("x_0_1", P4.slice(0, 384), g("x_0_1", N, 1, *sz[4], 128, hl))
        up(x11, P2.slice(320, 64))                          # slot E again: x_2_2 has consumed up(f/4)
        block("x_1_2", P2.slice(192, 192), P2.slice(128, 64))
        up(x01, P2.slice(0, 128))
        x02 = block("x_0_2", P2.slice(0, 320), g("x_0_2", N, 1, *sz[2], 64, hl))
        # ---- x_0_3 and the segmentation head (activation=None) at half resolution in space-to-depth form, then the
        # raw-reshape glue into the encoder clip
        t = conv_auto(x02, self.tail1, g("x_0_3.t", N, 1, *sz[2], 128, hl))
        h = conv_auto(t, self.tail2, g("x_0_3", N, 1, *sz[2], 128, hl))
        hs = getattr(self.head, "slab", None)
        if UNETPP_FUSE_GLUE and frames_out is None and enc_in.C in (4, 8) and ops.slab_runs_kx(h, hs):
            # the KX epilogue scatters the 12 space-to-depth channels straight into the encoder clip: no head tensor, no
            # glue kernel (0.23 ms and 0.4 GB per 32 clips); the clip's pad channels stay as allocated (zero)
            ops.conv_slab_forward(h, hs, None, act=L.ACT_NONE, s2d_clip=(enc_in, T))
            return enc_in
        out16 = conv_auto(h, self.head, g("head", N, 1, *sz[2], 16, hl), act=L.ACT_NONE)
        ops.frames_to_clip(out16, enc_in, T, frames_out, s2d=True)
        return enc_in


# ======================================================================================== I3D
I3D_MIXED = [
    ("Mixed_3b", 192, [64, 96, 128, 16, 32, 32]),
    ("Mixed_3c", 256, [128, 128, 192, 32, 96, 64]),
    ("Mixed_4b", 480, [192, 96, 208, 16, 48, 64]),
    ("Mixed_4c", 512, [160, 112, 224, 24, 64, 64]),
    ("Mixed_4d", 512, [128, 128, 256, 24, 64, 64]),
    ("Mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
    ("Mixed_4f", 528, [256, 160, 320, 32, 128, 128]),
    ("Mixed_5b", 832, [256, 160, 320, 32, 128, 128]),
    ("Mixed_5c", 832, [384, 192, 384, 48, 128, 128]),
]


class I3DExecutor:
    """InceptionI3d.extract_features (aux_code/models/i3d.py:336-340): Unit3D = TF-SAME conv + BN(eps 1e-3)
    + ReLU in one kernel; the four Inception branches write into slices of the block output."""

    def __init__(self, sd, device):
        self.device = device
        self.bufs = _Buffers(device)
        self.graphs = GraphCache()
        self.sd_w = {}
        self.specs = {}

        def unit(name, k, s=(1, 1, 1), cin_pad=None):
            w = sd[f"{name}.conv3d.weight"]
            if USE_SLAB and k == (3, 3, 3) and w.shape[1] >= (16 if PAD_SMALL_3X3 else 64):
                # SLAB feed: 64-channel K blocks (pad channels are zero).  Also for the 16..48-channel b2b branches:
                # up to 4x the MMA work, but through the slab feed instead of the instruction-bound GATHER feed
                cin_pad = -(-w.shape[1] // 64) * 64
            self.specs[name] = (w, _bn(sd, f"{name}.bn", 1e-3), k, s, cin_pad)

        unit("Conv3d_1a_7x7", (7, 7, 7), (2, 2, 2), 8)
        unit("Conv3d_2b_1x1", (1, 1, 1))
        unit("Conv3d_2c_3x3", (3, 3, 3))
        for name, _, _ in I3D_MIXED:
            for br, k in (("b0", 1), ("b1a", 1), ("b1b", 3), ("b2a", 1), ("b2b", 3), ("b3b", 1)):
                unit(f"{name}.{br}", (k, k, k))
        self.packed = {}
        self.in_layout = {}   # unit name -> (physical->logical input channel list | None, physical input channels)
        self.layout = {}      # block name -> (logical channels, physical->logical list | None) of its output buffer
        self.heads_slab = I3D_HEADS_SLAB
        if self.heads_slab:
            perm_in = None
            for i, (name, cin, oc) in enumerate(I3D_MIXED):
                for br in ("b0", "b1a", "b2a", "b3b"):   # the units that read the block input (or its max-pool)
                    self.in_layout[f"{name}.{br}"] = (perm_in, -(-cin // 64) * 64)
                total = oc[0] + oc[2] + oc[4] + oc[5]
                # physical [b1b | b2b | b3b | b0]: channel p of the buffer is logical channel perm[p]; the last block
                # keeps the reference's order (its output is the feature map)
                perm_in = None if i == len(I3D_MIXED) - 1 else list(range(oc[0], total)) + list(range(oc[0]))
                self.layout[name] = (total, perm_in)

    def _weight(self, name):
        """Unit3D weight [Cout, Cin_phys, kd, kh, kw] for the physical channel order / padding of its input buffer."""
        w = self.specs[name][0]
        perm, cphys = self.in_layout.get(name, (None, None))
        if cphys is None:
            return w
        wl = w.detach().float()
        if perm is not None:
            wl = wl.index_select(1, torch.as_tensor(perm, dtype=torch.long, device=wl.device))
        out = wl.new_zeros((wl.shape[0], cphys) + tuple(wl.shape[2:]))
        out[:, :wl.shape[1]] = wl
        return out

    def tap(self, name):
        """Block output in the reference's channel order (fp32 [N,C,D,H,W]) - for the per-layer parity tests."""
        buf = self.bufs.find(name)
        if buf is None or name not in self.layout:
            return None if buf is None else buf.to_ncdhw()
        total, perm = self.layout[name]
        t = buf.slice(0, total).to_ncdhw()
        if perm is None:
            return t
        inv = torch.empty(total, dtype=torch.long)
        inv[torch.as_tensor(perm)] = torch.arange(total)
        return t.index_select(1, inv.to(t.device))

    def _packed_unit(self, name, x):
        """PackedConv of a Unit3D for this input extent (1x1x1 units: no padding involved)."""
        w, bn, k, s, cin_pad = self.specs[name]
        key = (name, (0, 0, 0))
        pc = self.packed.get(key)
        if pc is None:
            assert k == (1, 1, 1)
            pc = PackedConv(self._weight(name), None, bn, stride=s, pad_front=(0, 0, 0), cin_pad=cin_pad, device=self.device, n_align=16)
            pc.cin = w.shape[1]
            pc.slab = None
            self.packed[key] = pc
        return pc

    def _conv(self, name, x, y):
        """Unit3D.forward (i3d.py:89-120): SAME front pads depend on the input extent."""
        w, bn, k, s, cin_pad = self.specs[name]
        pads = [same_pad(sz, kk, ss) for sz, kk, ss in zip((x.D, x.H, x.W), k, s)]
        pf = tuple(p[0] for p in pads)
        key = (name, pf)
        pc = self.packed.get(key)
        if pc is None:
            heads = name.rsplit(".", 1)[-1] in ("b0", "b1a", "b2a") and MERGE_1X1   # merged: general kernel
            pc = PackedConv(self._weight(name), None, bn, stride=s, pad_front=pf, cin_pad=cin_pad, device=self.device,
                            n_align=16 if (k == (1, 1, 1) and (heads or w.shape[0] % 8)) else 32)
            pc.cin = w.shape[1]
            self.packed[key] = pc
            if cin_pad == 8:
                self.packed[key + ("slab",)] = stem3d(pc)
            else:
                pc.slab = slab3x3(pc)
        if cin_pad == 8:
            return stem_conv(x, pc, self.packed[key + ("slab",)], y)
        return conv_auto(x, pc, y)

    @staticmethod
    def _same_out(x, s):
        return tuple(-(-sz // ss) for sz, ss in zip((x.D, x.H, x.W), s))

    def _pool(self, name, x, k, s):
        """MaxPool3dSamePadding.forward (i3d.py:21-45): zero padding, then max."""
        od, oh, ow = self._same_out(x, s)
        y = self.bufs.get(name, x.N, od, oh, ow, x.C)
        pf = tuple(same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((x.D, x.H, x.W), k, s))
        return ops.maxpool(x, y, k, s, pf, zero_pad=True)

    def _unit(self, name, x, cout=None, out=None):
        _, _, k, s, _ = self.specs[name]
        od, oh, ow = self._same_out(x, s)
        if out is None:
            out = self.bufs.get(name, x.N, od, oh, ow, cout)
        return self._conv(name, x, out)

    def _mixed_packed(self, name, x, oc):
        """InceptionModule.forward (i3d.py:142-149) with the heads as ONE single-destination slab convolution: the block
        buffer is physically [b1b | b2b | b3b | b0 | b1a (c1) | b2a (c2)], the heads write the contiguous range
        [b0 | b1a | b2a] (pad rows of b1a / b2a have zero weights and zero bias: they store zeros), the 3x3x3 branches
        read their padded inputs from the tail of the same buffer.  Returns the view the next layer reads: the first
        ceil64(total) physical channels (what lies beyond `total` meets zero weights)."""
        total = oc[0] + oc[2] + oc[4] + oc[5]
        c1, c2 = self.specs[f"{name}.b1b"][4], self.specs[f"{name}.b2b"][4]
        rows = oc[0] + c1 + c2
        y = self.bufs.get(name, x.N, x.D, x.H, x.W, total + c1 + c2)
        key = (name, "heads_slab")
        pc = self.packed.get(key)
        if pc is None:
            ws = [self._weight(f"{name}.{br}").detach().float() for br in ("b0", "b1a", "b2a")]
            begins = (0, oc[0], oc[0] + c1)
            w = ws[0].new_zeros((rows,) + tuple(ws[0].shape[1:]))
            g, b, m, v = (torch.ones(rows), torch.zeros(rows), torch.zeros(rows), torch.ones(rows))
            g, b, m, v = (t.to(ws[0].device) for t in (g, b, m, v))
            for br, wi, r0 in zip(("b0", "b1a", "b2a"), ws, begins):
                gi, bi, mi, vi, eps = self.specs[f"{name}.{br}"][1]
                n = wi.shape[0]
                w[r0:r0 + n] = wi
                g[r0:r0 + n], b[r0:r0 + n], m[r0:r0 + n], v[r0:r0 + n] = (t.detach().float() for t in (gi, bi, mi, vi))
            pc = PackedConv(w, None, (g, b, m, v, eps), pad_front=(0, 0, 0), cin_pad=w.shape[1], device=self.device, n_align=32)
            pc.cin = self.specs[f"{name}.b0"][0].shape[1]
            pc.slab = slab3x3(pc)
            assert pc.slab is not None, (name, pc.cout, pc.cout_pad, pc.n_tile)
            self.packed[key] = pc
        conv_auto(x, pc, y.slice(total - oc[0], rows))
        self._unit(f"{name}.b1b", y.slice(total, c1), out=y.slice(0, oc[2]))
        self._unit(f"{name}.b2b", y.slice(total + c1, c2), out=y.slice(oc[2], oc[4]))
        t3 = self._pool(f"{name}.b3a", x, (3, 3, 3), (1, 1, 1))
        self._unit(f"{name}.b3b", t3, out=y.slice(oc[2] + oc[4], oc[5]))
        return y.slice(0, -(-total // 64) * 64)

    def _mixed(self, name, x, oc):
        if self.heads_slab and self.layout[name][1] is not None:
            return self._mixed_packed(name, x, oc)
        total = oc[0] + oc[2] + oc[4] + oc[5]
        y = self.bufs.get(name, x.N, x.D, x.H, x.W, total)
        # b1a's output is stored with the channel padding b1b's feed wants (pad channels zero, never written)
        c1 = self.specs[f"{name}.b1b"][4] or oc[1]
        t1 = self.bufs.get(f"{name}.b1a", x.N, x.D, x.H, x.W, c1, zero=True)
        c2 = self.specs[f"{name}.b2b"][4] or oc[3]
        t2 = self.bufs.get(f"{name}.b2a", x.N, x.D, x.H, x.W, c2, zero=True)

        def b1():
            self._unit(f"{name}.b1b", t1, out=y.slice(oc[0], oc[2]))

        def b2():
            self._unit(f"{name}.b2b", t2, out=y.slice(oc[0] + oc[2], oc[4]))

        def b3():
            t3 = self._pool(f"{name}.b3a", x, (3, 3, 3), (1, 1, 1))
            self._unit(f"{name}.b3b", t3, out=y.slice(oc[0] + oc[2] + oc[4], oc[5]))

        # (running the four branches on side streams was measured on B200 in round 1 with eager launches: no gain)
        heads = [y.slice(0, oc[0]), t1.slice(0, oc[1]), t2.slice(0, oc[3])]
        fork = I3D_BRANCH_STREAMS and x.buf.is_cuda
        if fork:
            main = torch.cuda.current_stream()
            if not hasattr(self, "_side"):
                self._side = (torch.cuda.Stream(device=x.buf.device), torch.cuda.Stream(device=x.buf.device))
            ev_x = torch.cuda.Event()
            ev_x.record(main)
            self._side[0].wait_event(ev_x)
            with torch.cuda.stream(self._side[0]):
                b3()                              # max-pool + 1x1x1: depends on the block input only
        if MERGE_1X1:
            # the three 1x1x1 convolutions that read x (i3d.py:144-148) as ONE GEMM over stacked weight rows: x crosses
            # L2->SM once instead of three times and two launches disappear per block
            key = (name, "heads")
            pcm = self.packed.get(key)
            if pcm is None:
                pcm = ops.PackedConv.concat([self._packed_unit(f"{name}.{br}", x) for br in ("b0", "b1a", "b2a")])
                self.packed[key] = pcm
            ops.conv_forward(x, pcm, heads)
        else:
            for br, out in zip(("b0", "b1a", "b2a"), heads):
                self._unit(f"{name}.{br}", x, out=out)
        if fork:
            ev_h = torch.cuda.Event()
            ev_h.record(main)
            self._side[1].wait_event(ev_h)
            with torch.cuda.stream(self._side[1]):
                b2()
            b1()
            main.wait_stream(self._side[0])
            main.wait_stream(self._side[1])
        else:
            b1(); b2(); b3()
        return y

    def run_trunk(self, enc_in):
        """enc_in: [B,T,H,W,4|8] (3 real channels) -> Mixed_5c feature map [B,T/8,H/32,W/32,1024]."""
        x = self._unit("Conv3d_1a_7x7", enc_in, 64)
        x = self._pool("MaxPool3d_2a_3x3", x, (1, 3, 3), (1, 2, 2))
        x = self._unit("Conv3d_2b_1x1", x, 64)
        x = self._unit("Conv3d_2c_3x3", x, 192)
        x = self._pool("MaxPool3d_3a_3x3", x, (1, 3, 3), (1, 2, 2))
        mixed = {n: oc for n, _, oc in I3D_MIXED}
        for n in ("Mixed_3b", "Mixed_3c"):
            x = self._mixed(n, x, mixed[n])
        x = self._pool("MaxPool3d_4a_3x3", x, (3, 3, 3), (2, 2, 2))
        for n in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"):
            x = self._mixed(n, x, mixed[n])
        x = self._pool("MaxPool3d_5a_2x2", x, (2, 2, 2), (2, 2, 2))
        for n in ("Mixed_5b", "Mixed_5c"):
            x = self._mixed(n, x, mixed[n])
        return x

    def run(self, enc_in):
        """enc_in: [B,T,H,W,8] -> fp32 features [B, T', 1024] (AvgPool3d([2,7,7], stride 1), i3d.py:293-294,340)."""
        x = self.run_trunk(enc_in)
        if x.H != 7 or x.W != 7 or x.D < 2:
            raise RuntimeError(f"InceptionI3d.extract_features: AvgPool3d([2,7,7]) needs a (>=2,7,7) map, got "
                               f"({x.D},{x.H},{x.W}); use 224x224 clips of >=16 frames (i3d.py:293-294)")
        return ops.avgpool_features(x, 2)


# ==================================================================================== I3Res50
I3RES50_LAYERS = [(64, 3, 1, [1, 1, 1]), (128, 4, 2, [1, 0, 1, 0]), (256, 6, 2, [1, 0, 1, 0, 1, 0]), (512, 3, 2, [0, 1, 0])]


class I3Res50Executor:
    """I3Res50.extract_features (aux_code/models/large_i3d.py:249-263).  bn3 + residual add + ReLU
    (large_i3d.py:72-79) run in the epilogue of conv3."""

    def __init__(self, sd, device, prefix="i3d."):
        self.device = device
        self.bufs = _Buffers(device)
        self.graphs = GraphCache()
        P = prefix
        def mk(wk, bnk, stride, pad, cin_pad=None):
            pc = PackedConv(sd[P + wk], None, _bn(sd, P + bnk, 1e-5), stride=stride, pad_front=pad, cin_pad=cin_pad,
                            device=device, n_align=32)
            pc.slab = slab3x3(pc)     # the stride-1 (1,3,3) conv2 of most Bottlenecks
            return pc
        self.conv1 = mk("conv1.weight", "bn1", (2, 2, 2), (2, 3, 3), 8)
        self.conv1_slab = stem3d(self.conv1)
        self.blocks = []
        for li, (planes, nblocks, stride, tcs) in enumerate(I3RES50_LAYERS, 1):
            for b in range(nblocks):
                p = f"layer{li}.{b}"
                st = stride if b == 0 else 1
                blk = {
                    "c1": mk(f"{p}.conv1.weight", f"{p}.bn1", (1, 1, 1), (tcs[b], 0, 0)),
                    "c2": mk(f"{p}.conv2.weight", f"{p}.bn2", (1, st, st), (0, 1, 1)),
                    "c3": mk(f"{p}.conv3.weight", f"{p}.bn3", (1, 1, 1), (0, 0, 0)),
                    "ds": mk(f"{p}.downsample.0.weight", f"{p}.downsample.1", (1, st, st), (0, 0, 0)) if b == 0 else None,
                    "name": p, "pool_after": (li == 1 and b == nblocks - 1),
                }
                self.blocks.append(blk)

    def _apply(self, name, pc, x, res=None, act=L.ACT_RELU):
        od, oh, ow = pc.out_extent((x.D, x.H, x.W))
        y = self.bufs.get(name, x.N, od, oh, ow, pc.cout)
        return conv_auto(x, pc, y, res=res, act=act)

    def run(self, enc_in):
        """enc_in: [B,T,H,W,4|8] -> fp32 features [B, 1, 2048]."""
        g = self.bufs.get
        od, oh, ow = self.conv1.out_extent((enc_in.D, enc_in.H, enc_in.W))
        x = stem_conv(enc_in, self.conv1, self.conv1_slab, g("conv1", enc_in.N, od, oh, ow, self.conv1.cout))
        y = g("maxpool1", x.N, (x.D - 2) // 2 + 1, (x.H - 3) // 2 + 1, (x.W - 3) // 2 + 1, x.C)
        x = ops.maxpool(x, y, (2, 3, 3), (2, 2, 2))
        for blk in self.blocks:
            n = blk["name"]
            o = self._apply(n + ".c1", blk["c1"], x)
            o = self._apply(n + ".c2", blk["c2"], o)
            res = x if blk["ds"] is None else self._apply(n + ".ds", blk["ds"], x, act=L.ACT_NONE)
            x = self._apply(n + ".c3", blk["c3"], o, res=res)
            if blk["pool_after"]:
                y = g("maxpool2", x.N, (x.D - 2) // 2 + 1, x.H, x.W, x.C)
                x = ops.maxpool(x, y, (2, 1, 1), (2, 1, 1))
        return ops.avgpool_features(x, 0)


# ===================================================================================== R3D-18
class R3D18Executor:
    """wrapper_r3d_18.forward (aux_code/model_loaders.py:210-213) over torchvision's VideoResNet
    (video/resnet.py:251-263): returns (pred, feature)."""

    def __init__(self, sd, device):
        self.device = device
        self.bufs = _Buffers(device)
        self.graphs = GraphCache()
        def mk(wk, bnk, stride, pad, cin_pad=None):
            pc = PackedConv(sd[wk], None, _bn(sd, bnk, 1e-5), stride=stride, pad_front=pad, cin_pad=cin_pad, device=device,
                            n_align=32)
            pc.slab = slab3x3(pc)     # stride-1 3x3x3 convolutions (used when no residual is added in the epilogue)
            return pc
        self.stem = mk("backbone.stem.0.weight", "backbone.stem.1", (1, 2, 2), (1, 3, 3), 8)
        self.stem_slab = stem3d(self.stem)
        self.blocks = []
        for li in range(1, 5):
            for b in range(2):
                p = f"backbone.layer{li}.{b}"
                st = 2 if (b == 0 and li > 1) else 1
                self.blocks.append({
                    "c1": mk(f"{p}.conv1.0.weight", f"{p}.conv1.1", (st, st, st), (1, 1, 1)),
                    "c2": mk(f"{p}.conv2.0.weight", f"{p}.conv2.1", (1, 1, 1), (1, 1, 1)),
                    "ds": mk(f"{p}.downsample.0.weight", f"{p}.downsample.1", (st, st, st), (0, 0, 0))
                    if (b == 0 and li > 1) else None,
                    "name": p})
        self.fc = PackedConv(sd["fc.weight"], sd["fc.bias"], None, device=device)

    def _apply(self, name, pc, x, res=None, act=L.ACT_RELU):
        od, oh, ow = pc.out_extent((x.D, x.H, x.W))
        y = self.bufs.get(name, x.N, od, oh, ow, pc.cout)
        return conv_auto(x, pc, y, res=res, act=act)

    def run(self, enc_in):
        od, oh, ow = self.stem.out_extent((enc_in.D, enc_in.H, enc_in.W))
        x = stem_conv(enc_in, self.stem, self.stem_slab, self.bufs.get("stem", enc_in.N, od, oh, ow, self.stem.cout))
        for blk in self.blocks:
            n = blk["name"]
            o = self._apply(n + ".c1", blk["c1"], x)
            res = x if blk["ds"] is None else self._apply(n + ".ds", blk["ds"], x, act=L.ACT_NONE)
            x = self._apply(n + ".c2", blk["c2"], o, res=res)
        feat = ops.avgpool_features(x, 0)  # [B,1,512] fp32
        fin = self.bufs.get("fc_in", x.N, 1, 1, 1, 512)
        ops.nchw_to_cl(feat.reshape(x.N, 512, 1, 1, 1), fin)  # fp32 -> bf16 operand of the head GEMM (own kernel, no ATen)
        pred = self.bufs.get("fc_out", x.N, 1, 1, 1, self.fc.cout, dtype=torch.float32)
        ops.conv_forward(fin, self.fc, pred, act=L.ACT_NONE, y_fp32=True)
        return pred.buf.reshape(x.N, self.fc.cout), feat.reshape(x.N, 512)
