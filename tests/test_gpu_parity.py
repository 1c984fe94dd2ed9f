"""GPU parity tests (run on the B200 box with `-m gpu`): every C-ABI operator against the matching
torch fp32 op, and the whole hot path (uint8 frames -> preprocessing -> anonymizer -> encoder ->
feature rows) against the fp32 oracle and the golden vectors of the unmodified reference.

Gate (BASELINE.json north_star): per-snippet feature cosine >= 0.9995 and max abs error <= 2e-2 for
the bf16 pipeline versus the fp32 reference; snippet (temporal) index and crop ordering bit-exact.
"""
import os
import sys

import numpy as np
import pytest
import torch

import _cases
from oracle import models as M
from oracle import preprocess as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))

pytestmark = pytest.mark.gpu

COS_GATE, ABS_GATE = 0.9995, 2e-2
# Scale-free companions of the absolute gate (the absolute gate depends on the synthetic feature
# scale).  REL_L2 = ||got - ref|| / ||ref|| (cos >= 0.9995 <=> <= 0.0316 for unbiased noise); REL_MAX = worst element
# over the largest feature; TAP_REL = relative RMS error of every intermediate activation the executors keep, against
# the oracle's taps of the same layer.  Measured on B200 (profiles/r2_per_layer_errors.txt): bf16 storage noise starts at
# 0.006-0.008 after the first convolution and grows SMOOTHLY with depth (x1.0-1.7 per layer; 0.024 at the UNet
# bottleneck, 0.040 at Mixed_5c, 0.053 at I3Res50.layer4.0 after ~65 stacked convolutions).  A kernel bug shows up at
# ITS layer at the size of the signal (0.1-1.0) and as a jump: gate = absolute cap + bounded growth per layer.
REL_L2_GATE = 0.0316
REL_MAX_GATE = {"i3d": 0.10, "largei3d": 0.03, "r3d_18": 0.03}
TAP_REL_GATE = 0.08
TAP_GROWTH_GATE = (2.0, 0.004)     # err[layer] <= 2.0 * max(err of earlier layers) + 0.004


def _modules(name, stress=False):
    from aux_code.model_loaders import load_fa_model, load_ft_model
    arch = _cases.CASES[name][0]
    sd_fa, sd_ft = _cases.case_weights(name, stress)
    fa = load_fa_model(arch=_cases.fa_arch(name))
    ft = load_ft_model(arch=arch, num_classes=102)
    fa.load_state_dict(sd_fa, strict=True)
    ft.load_state_dict(sd_ft, strict=True)
    return fa.cuda().eval(), ft.cuda().eval()


@pytest.mark.parametrize("group", ["flat", "gather", "multi", "ops", "prep", "slab3", "slabpair", "slabstream", "streampair",
                                   "slab1x1", "slabstem", "slabup", "slabkx", "fuzz"])
def test_operator_battery(group):
    """tests/gpu_diag.py: each operator vs torch fp32 on bf16-rounded operands (conv tolerance 2e-2 of the
    output range, i.e. bf16 output rounding; pooling / layout / PIL preprocessing bit-exact)."""
    import gpu_diag
    gpu_diag.RESULTS.clear()
    getattr(gpu_diag, "group_" + group)()
    torch.cuda.synchronize()
    failed = [n for n, ok in gpu_diag.RESULTS if not ok]
    assert not failed, failed
    assert len(gpu_diag.RESULTS) >= 4


def _pipeline_features(ext, name, which, hw):
    from tedspad_b200.extraction import crop_boxes
    clip = _cases.case_clip(name, which)
    (ch, cw), boxes = crop_boxes(hw[0], hw[1])
    desc = np.zeros((16, 4), dtype=np.int32)
    desc[:, 0] = np.arange(16)
    desc[:, 1], desc[:, 2] = boxes[0][0], boxes[0][1]
    f = ext.features_of_clips(torch.from_numpy(np.ascontiguousarray(clip)).cuda(), desc, (ch, cw))
    torch.cuda.synchronize()
    return clip, f.reshape(-1).float().cpu()


def _rel_l2(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    return float((got - ref).norm() / ref.norm())


def _depth_to_space(cl, C):
    """CLTensor [N,1,h,w,>=4C] in space-to-depth form (channel (2a+b)*C + c) -> torch [N,C,2h,2w]."""
    v = cl.to_ncdhw()[:, :4 * C, 0].cpu()                                 # [N,4C,h,w]
    N, _, h, w = v.shape
    return v.reshape(N, 2, 2, C, h, w).permute(0, 3, 4, 1, 5, 2).reshape(N, C, 2 * h, 2 * w)


class _Dense:
    """Adapter so that a plain [N,C,H,W] tensor passes through _tap_errors like a CLTensor."""

    def __init__(self, t):
        self.t = t

    def to_ncdhw(self):
        return self.t.unsqueeze(2)


class _Dense5:
    """The same for a tensor that already is [N,C,D,H,W]."""

    def __init__(self, t):
        self.t = t

    def to_ncdhw(self):
        return self.t


def _tap_errors(fa, ft, arch, taps_fa, taps_ft, B=1):
    """{layer: relative RMS error} of the activation buffers the executors keep after a run vs the oracle's taps."""
    from tedspad_b200.engine import UNetExecutor
    dev = next(fa.parameters()).device
    ex_fa = fa.executor(dev)
    enc_mod = ft.i3d if hasattr(ft, "i3d") else ft
    ex_ft = enc_mod._exec(torch.empty(1, device=dev))
    pairs = []   # (label, CLTensor, oracle tensor [N,C,(D,)H,W])
    if isinstance(ex_fa, UNetExecutor):
        lv = UNetExecutor.LEVELS
        for i, p in enumerate(lv):
            pairs.append((f"unet:{p}.0", ex_fa.bufs.find(f"t{i}"), taps_fa[f"{p}.0"]))
            b = ex_fa.bufs.find(f"cat{i}") if i < 4 else ex_fa.bufs.find("x5")
            pairs.append((f"unet:{p}.3", b.slice(0, taps_fa[f"{p}.3"].shape[1]) if i < 4 else b, taps_fa[f"{p}.3"]))
        for j in range(4):
            p = f"up{j + 1}.conv.double_conv"
            pairs.append((f"unet:{p}.0", ex_fa.bufs.find(f"u{j}a"), taps_fa[f"{p}.0"]))
            if j < 3:
                pairs.append((f"unet:{p}.3", ex_fa.bufs.find(f"u{j}"), taps_fa[f"{p}.3"]))
    else:   # UNet++: encoder stages and every decoder block output, where the executor keeps them (engine.UNetPPExecutor)
        P2, P4, P8 = (ex_fa.bufs.find(n) for n in ("P2", "P4", "P8"))
        where = {"encoder.conv1": P2.slice(256, 64), "encoder.layer1.1": P4.slice(320, 64), "encoder.layer2.1": P8.slice(256, 128),
                 "encoder.layer3.1": ex_fa.bufs.find("f16"), "decoder.blocks.x_0_0.conv2": ex_fa.bufs.find("x_0_0"),
                 "decoder.blocks.x_1_1.conv2": P4.slice(256, 64), "decoder.blocks.x_2_2.conv2": P2.slice(192, 64),
                 "decoder.blocks.x_0_1.conv2": ex_fa.bufs.find("x_0_1"), "decoder.blocks.x_1_2.conv2": P2.slice(128, 64),
                 "decoder.blocks.x_0_2.conv2": ex_fa.bufs.find("x_0_2"),
                 "decoder.blocks.x_0_3.conv2": _Dense(_depth_to_space(ex_fa.bufs.find("x_0_3"), 32)),
                 "out": _Dense(_depth_to_space(ex_fa.bufs.find("head"), 3)) if ex_fa.bufs.find("head") is not None else None}
        where = {k: b for k, b in where.items() if b is not None}   # (the head tensor does not exist when its glue is fused)
        for k, b in where.items():
            pairs.append((f"unetpp:{k}", b, taps_fa[k]))
    if arch == "i3d":
        for n in ["Conv3d_1a_7x7", "Conv3d_2b_1x1", "Conv3d_2c_3x3"] + [m[0] for m in M.I3D_MIXED]:
            t = ex_ft.tap(n)   # reference channel order (the Mixed buffers are stored permuted, engine.I3D_HEADS_SLAB)
            pairs.append((f"i3d:{n}", None if t is None else _Dense5(t), taps_ft[n]))
    elif arch == "largei3d":
        pairs.append(("i3res50:conv1", ex_ft.bufs.find("conv1"), taps_ft["conv1"]))
        for blk in ex_ft.blocks:
            # (layer1's last block is followed by the temporal max-pool in the oracle tap: compare the others)
            if not blk["pool_after"]:
                pairs.append((f"i3res50:{blk['name']}", ex_ft.bufs.find(blk["name"] + ".c3"), taps_ft["i3d." + blk["name"]]))
    else:
        pairs.append(("r3d:stem", ex_ft.bufs.find("stem"), taps_ft["stem"]))
        for blk in ex_ft.blocks:
            pairs.append((f"r3d:{blk['name']}", ex_ft.bufs.find(blk["name"] + ".c2"), taps_ft[blk["name"]]))
    out = {}
    for label, buf, ref in pairs:
        assert buf is not None, label
        got = buf.to_ncdhw().cpu()
        got = got[:ref.shape[0]]
        if ref.dim() == 4:
            got = got[:, :, 0]
        got = got[:, :ref.shape[1]]
        assert got.shape == ref.shape, (label, got.shape, ref.shape)
        out[label] = _rel_l2(got, ref)
    return out


def _torch_autocast_bf16_features(name, x_ref, stress=False):
    """Yardstick: the oracle's own functional model run by stock PyTorch on the GPU under
    torch.autocast(bfloat16) (cuDNN bf16 kernels) - how far *any* bf16 evaluation of this network is from fp32."""
    arch = _cases.CASES[name][0]
    sd_fa, sd_ft = _cases.case_weights(name, stress)
    sd_fa = {k: v.cuda() for k, v in sd_fa.items()}
    sd_ft = {k: v.cuda() for k, v in sd_ft.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        enc = M.anonymize_and_reshape(sd_fa, x_ref.cuda().unsqueeze(0))
        return M.encoder_features(arch, sd_ft, enc)[0].float().cpu()


@pytest.mark.parametrize("name", list(_cases.CASES))
def test_hot_path_parity(name):
    """uint8 frames -> preprocessing -> UNet -> glue -> encoder -> feature row, against the fp32 oracle, the
    golden features of the unmodified reference, and the torch-autocast-bf16 yardstick."""
    from tedspad_b200.extraction import SnippetExtractor
    arch, hw, reso, _, _ = _cases.CASES[name]
    fa, ft = _modules(name)
    ext = SnippetExtractor(fa, ft, reso=reso, batch_clips=2)
    G = _cases.golden()
    feats, refs = {}, {}
    for which in ("test", "control"):
        clip, feats[which] = _pipeline_features(ext, name, which, hw)
        x_ref, enc_ref, refs[which] = _cases.oracle_features(name, clip)
        if which == "test":
            # anonymized clip as the encoder sees it (bf16, scattered by the raw-reshape glue)
            enc = ext._enc_in(1).to_ncdhw()[:, :3].cpu()
            e = (enc - enc_ref).abs()
            print(f"\n{name}: anonymized clip err rms={e.pow(2).mean().sqrt():.4f} max={e.max():.4f}")
            assert e.pow(2).mean().sqrt() < 0.012 and e.max() < 0.1
            yard = _cases.parity_metrics(_torch_autocast_bf16_features(name, x_ref), refs[which])
    m = _cases.parity_metrics(feats["test"], refs["test"])
    # golden vectors of the unmodified reference exist for every case whose anonymizer the reference can build here
    # (arch='unet'); the unet++ case is checked against the restated oracle only (smp not installable: parity unpinned)
    has_golden = f"{name}/features" in G.files
    mg = _cases.parity_metrics(feats["test"], torch.from_numpy(G[f"{name}/features"])) if has_golden else m
    ctrl = float(torch.nn.functional.cosine_similarity(refs["test"], refs["control"], dim=0))
    dcos = float(torch.nn.functional.cosine_similarity(feats["test"] - feats["control"], refs["test"] - refs["control"], dim=0))
    print(f"{name}: vs oracle cos={m['cos']:.6f} max_abs={m['max_abs']:.4f} (|f|max {m['ref_max']:.3f}); "
          f"vs golden(reference) cos={mg['cos']:.6f} max_abs={mg['max_abs']:.4f}; "
          f"torch autocast bf16 yardstick cos={yard['cos']:.6f} max_abs={yard['max_abs']:.4f}; "
          f"control cos(different clips)={ctrl:.4f}; cos of clip-to-clip feature difference={dcos:.4f}")
    assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, m
    assert mg["cos"] >= COS_GATE and mg["max_abs"] <= ABS_GATE, mg
    # scale-free: relative L2 and worst element relative to the largest feature
    rl2, rmax = _rel_l2(feats["test"], refs["test"]), m["max_abs"] / m["ref_max"]
    print(f"{name}: rel_l2={rl2:.5f} (gate {REL_L2_GATE}) max_abs/|f|max={rmax:.4f} (gate {REL_MAX_GATE[arch]})")
    assert rl2 <= REL_L2_GATE and rmax <= REL_MAX_GATE[arch]
    # per-layer: every intermediate activation of the LAST run (the control clip) against the oracle's taps
    taps_fa, taps_ft = {}, {}
    sd_fa, sd_ft = _cases.case_weights(name)
    with torch.no_grad():
        x_c = torch.from_numpy(P.dali_val_augmentations(_cases.case_clip(name, "control"), reso))
        enc_c = M.anonymize_and_reshape(sd_fa, x_c.unsqueeze(0), taps=taps_fa)
        M.encoder_features(arch, sd_ft, enc_c, taps=taps_ft)
    errs = _tap_errors(fa, ft, arch, taps_fa, taps_ft)
    worst = max(errs, key=errs.get)
    print(f"{name}: per-layer relative RMS error: " + ", ".join(f"{k.split(':')[1]}={v:.4f}" for k, v in errs.items()))
    print(f"{name}: worst layer {worst} {errs[worst]:.4f} (gate {TAP_REL_GATE})")
    assert errs[worst] <= TAP_REL_GATE, (worst, errs[worst])
    seen = 0.0
    for k, v in errs.items():       # insertion order = network order
        assert seen == 0.0 or v <= TAP_GROWTH_GATE[0] * seen + TAP_GROWTH_GATE[1], (k, v, seen)
        seen = max(seen, v)
    assert ctrl < COS_GATE            # the gate can tell two clips apart ...
    assert dcos > 0.98                # ... and the response to changing the clip matches the reference's
    # no worse than the existing bf16 kernels on the same network
    assert (1 - m["cos"]) <= 1.25 * (1 - yard["cos"]) + 1e-6 and m["max_abs"] <= 1.25 * yard["max_abs"] + 1e-3


def test_stress_init_no_worse_than_torch_bf16():
    """Chaotic synthetic init (BN beta ~ 0: perturbation gain ~1.2 per layer, SURVEY 8c's first suggestion).
    No 16-bit evaluation stays within the absolute gate there; what is checked is that this pipeline deviates
    from fp32 no more than stock PyTorch bf16 autocast does."""
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_largei3d_224"
    arch, hw, reso, _, _ = _cases.CASES[name]
    fa, ft = _modules(name, stress=True)
    ext = SnippetExtractor(fa, ft, reso=reso, batch_clips=1)
    clip, feat = _pipeline_features(ext, name, "test", hw)
    x_ref, _, ref = _cases.oracle_features(name, clip, stress=True)
    m = _cases.parity_metrics(feat, ref)
    yard = _cases.parity_metrics(_torch_autocast_bf16_features(name, x_ref, stress=True), ref)
    print(f"\nstress init: ours cos={m['cos']:.5f} max_abs={m['max_abs']:.3f}; torch autocast bf16 cos={yard['cos']:.5f} "
          f"max_abs={yard['max_abs']:.3f}")
    # aggregate deviation (cosine, RMS) within 1.25x of cuDNN-bf16's; the single worst element of a chaotic network is
    # a noisy statistic (it moved 0.31 -> 0.39 between two equally valid accumulation orders), so it gets 1.5x
    rms = lambda a, b: float((a.double() - b.double()).pow(2).mean().sqrt())  # noqa: E731
    yard_feat = _torch_autocast_bf16_features(name, x_ref, stress=True)
    assert (1 - m["cos"]) <= 1.25 * (1 - yard["cos"]) + 1e-6
    assert rms(feat, ref) <= 1.25 * rms(yard_feat, ref) + 1e-4
    assert m["max_abs"] <= 1.5 * yard["max_abs"] + 1e-3


def test_drop_in_module_flow():
    """The loop body of feature_extraction/dali_extraction.py:168-179 run verbatim on the product modules
    (fp32 torch tensors at every module boundary, bare try/except attribute dispatch)."""
    name = "unet_largei3d_224"
    fa_model, ft_model = _modules(name)
    clip = _cases.case_clip(name)
    x_ref, enc_ref, f_ref = _cases.oracle_features(name, clip)
    inputs = x_ref.unsqueeze(0).cuda()                       # what val_augmentations hands over: [1,16,3,224,224]
    with torch.no_grad():
        ori_bs, ori_t, ori_c, ori_h, ori_w = inputs.permute(0, 2, 1, 3, 4).shape
        inputs = inputs.view(-1, inputs.shape[2], inputs.shape[3], inputs.shape[4])
        anon = fa_model(inputs)
        assert anon.shape == (16, 3, 224, 224) and anon.dtype == torch.float32 and anon.is_cuda
        inputs = anon.reshape(ori_bs, ori_t, ori_c, ori_h, ori_w)
        try:
            output = ft_model.extract_features(inputs)
        except:  # noqa: E722  (the reference's own dispatch idiom, dali_extraction.py:175-178)
            output = ft_model.i3d.extract_features(inputs)
        assert output.shape == (1, 2048, 1, 1, 1) and output.dtype == torch.float32
        row = output.squeeze().cpu().numpy()
    e = (anon.cpu() - enc_ref.reshape(16, 3, 224, 224)).abs()
    assert e.max() < 0.1
    m = _cases.parity_metrics(torch.from_numpy(row), f_ref)
    print(f"\ndrop-in flow: cos={m['cos']:.6f} max_abs={m['max_abs']:.4f}")
    assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, m


def test_r3d18_forward_signature():
    """BASELINE config 1: pred, feat = ft_model(x) for wrapper_r3d_18 (model_loaders.py:210-213)."""
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    clip = _cases.case_clip(name)
    x_ref, enc_ref, f_ref = _cases.oracle_features(name, clip)
    sd_ft = _cases.case_weights(name)[1]
    with torch.no_grad():
        pred_ref, feat_ref = M.r3d18_forward(sd_ft, enc_ref)
        pred, feat = ft(enc_ref.cuda())
    assert pred.shape == (1, 102) and feat.shape == (1, 512)
    m = _cases.parity_metrics(feat.cpu(), feat_ref)
    mp = _cases.parity_metrics(pred.cpu(), pred_ref)
    print(f"\nr3d18 forward: feat cos={m['cos']:.6f} max_abs={m['max_abs']:.4f}; pred cos={mp['cos']:.6f}")
    assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE and mp["cos"] >= 0.999


def _video(n_frames, h, w, seed):
    base = M.structured_clip_u8(seed, 16, h, w)
    reps = -(-n_frames // 16)
    g = torch.Generator().manual_seed(seed)
    vid = torch.from_numpy(base).repeat(reps, 1, 1, 1)[:n_frames].clone()
    vid += torch.randint(0, 8, (n_frames, 1, 1, 3), generator=g, dtype=torch.uint8) * 4  # make every frame distinct
    return vid


def test_extract_video_rows_and_order():
    """Per-video matrix: float64 [ceil(N/32), F]; row i == features of frames 32i + 2j (zero images past the
    end); batched execution == clip-at-a-time execution bit for bit (temporal ordering gate)."""
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    vid = _video(150, 120, 160, 5)
    a = SnippetExtractor(fa, ft, reso=(112, 112), batch_clips=4).extract_video(vid)
    b = SnippetExtractor(fa, ft, reso=(112, 112), batch_clips=1).extract_video(vid.cuda())
    assert a.dtype == np.float64 and a.shape == (5, 512)
    assert np.array_equal(a, b)
    # oracle for rows 1 and 4 (row 4 is the zero-padded tail: frames 128..148 then zeros)
    for r in (1, 4):
        idx = M.dali_snippet_frames(150)[r]
        clip = np.stack([vid[i].numpy() if i >= 0 else np.zeros((120, 160, 3), np.uint8) for i in idx])
        _, _, f_ref = _cases.oracle_features(name, clip)
        m = _cases.parity_metrics(torch.from_numpy(a[r]), f_ref)
        assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, (r, m)
    # a permutation of the snippets must show up as the same permutation of rows
    assert not np.array_equal(a[0], a[1])


def test_full_size_config_properties():
    """BASELINE.json configs[1] at full size: 32 clips of 16x224x224 from 512 decoded 240x320 frames, UNet + I3D, one
    batch.  Size-independent properties (the fp32 oracle needs > 1 s per clip): the batched run equals the
    clip-at-a-time run bit for bit (tiles of the stacked-row / CTA-pair kernels straddle frames and clips, results
    must not), reversing the clips reverses the rows bit for bit, rows of different clips differ, and two rows are
    checked against the oracle with the north-star gate."""
    from tedspad_b200.extraction import SnippetExtractor, crop_boxes
    name = "unet_i3d_224"
    fa, ft = _modules(name)
    B, T = 32, 16
    clips = [M.structured_clip_u8(500 + i, T, 240, 320) for i in range(B)]
    frames = torch.from_numpy(np.concatenate(clips)).cuda()                       # [512, 240, 320, 3]
    (ch, cw), boxes = crop_boxes(240, 320)
    desc = np.zeros((B * T, 4), dtype=np.int32)
    desc[:, 0] = np.arange(B * T)
    desc[:, 1], desc[:, 2] = boxes[0][0], boxes[0][1]
    big = SnippetExtractor(fa, ft, batch_clips=B)
    f32 = big.features_of_clips(frames, desc, (ch, cw)).reshape(B, -1).float().cpu()
    assert f32.shape == (B, 1024) and bool(torch.isfinite(f32).all())
    one = SnippetExtractor(fa, ft, batch_clips=1)
    d1 = desc[:T].copy()
    for b in range(B):
        fb = one.features_of_clips(frames[b * T:(b + 1) * T], d1, (ch, cw)).reshape(-1).float().cpu()
        assert torch.equal(fb, f32[b]), f"clip {b}: batched != clip-at-a-time"
    rev = desc.reshape(B, T, 4)[::-1].reshape(B * T, 4).copy()
    f_rev = big.features_of_clips(frames, rev, (ch, cw)).reshape(B, -1).float().cpu()
    assert torch.equal(f_rev, f32.flip(0))
    cos = torch.nn.functional.cosine_similarity(f32[:-1], f32[1:], dim=1)
    assert float(cos.max()) < 0.99999, "different clips must give different rows"
    for b in (0, 21):
        _, _, f_ref = _cases.oracle_features(name, clips[b])
        m = _cases.parity_metrics(f32[b], f_ref)
        assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, (b, m)


def test_multicrop_layout_and_order():
    """[n_snip, ncrops, F]; crop 4 (center) is the reference's single crop, bit-exact; crops 5..9 are the
    crops of the horizontally flipped frames (torchvision ten_crop order)."""
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    vid = _video(64, 120, 160, 6)
    one = SnippetExtractor(fa, ft, reso=(112, 112), batch_clips=2).extract_video(vid)
    ten = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=10, batch_clips=20).extract_video(vid)
    five = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=5, batch_clips=5).extract_video(vid)
    assert ten.shape == (2, 10, 512) and five.shape == (2, 5, 512) and one.shape == (2, 512)
    assert np.array_equal(ten[:, 4], one) and np.array_equal(five, ten[:, :5])
    flipped = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=10, batch_clips=20).extract_video(vid.flip(2))
    # crops 5..9 are by definition crops 0..4 of the flipped frames, so flipping the video swaps the halves
    swap = [5, 6, 7, 8, 9, 0, 1, 2, 3, 4]
    assert np.array_equal(flipped, ten[:, swap])


def test_shanghai_path_preprocessing_and_indexing():
    """ShanghaiTech path: frames 32i+2j+1, tail dropped, crop (int(.8H), int(.8H)), Pillow 8-bit bilinear
    (bit-exact), channel order untouched (BGR stays BGR)."""
    from tedspad_b200 import ops, _lib as L
    from tedspad_b200.extraction import SnippetExtractor, crop_boxes
    name = "unet_largei3d_224"
    fa, ft = _modules(name)
    vid = _video(70, 240, 428, 7)
    ext = SnippetExtractor(fa, ft, source="shanghai", batch_clips=2)
    assert ext.snippet_frames(70).tolist() == np.asarray(M.shanghai_snippet_frames(70)).tolist()
    (ch, cw), boxes = crop_boxes(240, 428, square_from_h=True)
    desc = torch.tensor([[i, boxes[0][0], boxes[0][1], 0] for i in (1, 3)], dtype=torch.int32).cuda()
    y = ops.CLTensor(2, 1, 224, 224, 8, device="cuda")
    f32 = torch.empty(2, 3, 224, 224, device="cuda")
    ops.preprocess(vid.cuda(), desc, (ch, cw), y, L.RESAMPLE_PIL_U8, f32)
    ref = np.stack([P.shanghai_augmentation(vid[i].numpy()) for i in (1, 3)])
    assert np.array_equal(f32.cpu().numpy(), ref)
    feats = ext.extract_video(vid)
    assert feats.shape == (2, 2048) and feats.dtype == np.float64


def test_sharded_extraction_equals_single(tmp_path):
    """Union of 2 shards' files == single-process files, bit for bit (SURVEY 8e acceptance test)."""
    from tedspad_b200.extraction import SnippetExtractor, extract_dataset
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    ext = SnippetExtractor(fa, ft, reso=(112, 112), batch_clips=4)
    vids = [(f"/data/v{i}.mp4", n, (lambda n=n, i=i: _video(n, 120, 160, 20 + i))) for i, n in enumerate([40, 96, 33, 70])]
    single = tmp_path / "single"
    extract_dataset(ext, vids, str(single), log=lambda *_: None)
    sharded = tmp_path / "sharded"
    w0 = extract_dataset(ext, vids, str(sharded), rank=0, world_size=2, log=lambda *_: None)
    w1 = extract_dataset(ext, vids, str(sharded), rank=1, world_size=2, log=lambda *_: None)
    assert len(w0) + len(w1) == 4 and not (set(w0) & set(w1))
    for i in range(4):
        assert np.array_equal(np.load(single / f"v{i}.npy"), np.load(sharded / f"v{i}.npy"))


def test_packed_videos_equal_per_video():
    """extract_videos packs snippets of several (short) videos into full batches, also across the copy-stream staging
    buffers and around a change of frame size: rows must equal the per-video extract_video bit for bit, in order."""
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    lens = [40, 96, 33, 0, 70, 150, 31]
    vids = [_video(n, 120, 160, 30 + i) if n else torch.zeros((0, 120, 160, 3), dtype=torch.uint8) for i, n in enumerate(lens)]
    vids.insert(3, _video(50, 100, 140, 77))            # a different frame size in the middle: forces a flush
    ext = SnippetExtractor(fa, ft, reso=(112, 112), batch_clips=4)
    packed = list(ext.extract_videos(v.pin_memory() if v.shape[0] else v for v in vids))
    assert [i for i, _ in packed] == list(range(len(vids)))
    for (i, got), v in zip(packed, vids):
        want = ext.extract_video(v) if v.shape[0] else np.zeros((0, 0))
        assert got.dtype == np.float64 and got.shape == want.shape, (i, got.shape, want.shape)
        assert np.array_equal(got, want), i
    ten = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=10, batch_clips=20)
    packed10 = dict(ten.extract_videos(vids[:3]))
    for i in range(3):
        assert np.array_equal(packed10[i], ten.extract_video(vids[i]))


def test_shanghai_feature_parity():
    """ShanghaiTech path end to end at the dataset's frame size (480x856, BGR as cv2 decodes): snippet rows against the
    oracle (shanghai_dl.augmentation -> anonymizer -> raw-reshape glue -> I3Res50.extract_features) under the
    north-star gate.  Frames 32i+2j+1, crop 384x384 at (48, 236), Pillow 8-bit resize."""
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_largei3d_224"
    fa, ft = _modules(name)
    vid = _video(70, 480, 856, 9)
    feats = SnippetExtractor(fa, ft, source="shanghai", batch_clips=2).extract_video(vid)
    assert feats.shape == (2, 2048) and feats.dtype == np.float64
    sd_fa, sd_ft = _cases.case_weights(name)
    idx = M.shanghai_snippet_frames(70)
    for r in (0, 1):
        x = torch.from_numpy(np.stack([P.shanghai_augmentation(vid[i].numpy()) for i in idx[r]]))
        with torch.no_grad():
            f_ref = M.encoder_features("largei3d", sd_ft, M.anonymize_and_reshape(sd_fa, x.unsqueeze(0)))[0]
        m = _cases.parity_metrics(torch.from_numpy(feats[r]), f_ref)
        print(f"\nshanghai row {r}: cos={m['cos']:.6f} max_abs={m['max_abs']:.4f} rel_l2={_rel_l2(torch.from_numpy(feats[r]), f_ref):.5f}")
        assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, (r, m)
    assert not np.array_equal(feats[0], feats[1])


def test_multicrop_feature_parity_corner_and_flipped():
    """10-crop path: a corner crop (1 = top right) and a crop of the flipped frame (7 = bottom left of the h-flipped
    frame) as FEATURES against the oracle run on that crop (r1 only checked boxes and crop 4)."""
    from tedspad_b200.extraction import SnippetExtractor, crop_boxes
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    vid = _video(32, 120, 160, 11)
    ten = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=10, batch_clips=10).extract_video(vid)
    assert ten.shape == (1, 10, 512)
    (ch, cw), boxes = crop_boxes(120, 160, 10)
    sd_fa, sd_ft = _cases.case_weights(name)
    clip = vid[M.dali_snippet_frames(32)[0]].numpy()
    for ci in (1, 7, 4):
        x = torch.from_numpy(P.dali_crop_augmentations(clip, boxes[ci], (ch, cw), (112, 112)))
        with torch.no_grad():
            f_ref = M.encoder_features("r3d_18", sd_ft, M.anonymize_and_reshape(sd_fa, x.unsqueeze(0)))[0]
        m = _cases.parity_metrics(torch.from_numpy(ten[0, ci]), f_ref)
        print(f"\ncrop {ci} {boxes[ci]}: cos={m['cos']:.6f} max_abs={m['max_abs']:.4f}")
        assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, (ci, m)
    assert not np.array_equal(ten[0, 1], ten[0, 7])


def test_saved_files_load_like_the_mgfn_consumer(tmp_path):
    """SURVEY a-12: files written by extract_dataset, loaded exactly as anomaly_detection_mgfn's Dataset.__getitem__
    does (dataset.py:53-55 np.load -> float32; :70-71 / :87-89 expand_dims(axis=1), transpose to [ncrops,T,F];
    process_feat to 32 segments; magnitude appended) for the [T,F] and the [T,10,F] layouts."""
    from oracle import consumer as C
    from tedspad_b200.extraction import SnippetExtractor, extract_dataset
    name = "unet_r3d18_112"
    fa, ft = _modules(name)
    vids = [(f"/data/Abuse{i:03d}_x264.mp4", n, (lambda n=n, i=i: _video(n, 120, 160, 40 + i))) for i, n in enumerate([100, 33])]
    for ncrops, folder in ((1, "ucf_features_ours"), (10, "ucf_features_ours_10crop")):
        ext = SnippetExtractor(fa, ft, reso=(112, 112), ncrops=ncrops, batch_clips=10)
        out = tmp_path / folder
        written = extract_dataset(ext, vids, str(out), log=lambda *_: None)
        assert sorted(os.path.basename(w) for w in written) == ["Abuse000_x264.npy", "Abuse001_x264.npy"]
        for (path, n, _), T in zip(vids, (4, 2)):
            f = str(out / (os.path.basename(path).replace(".mp4", "") + ".npy"))
            raw = np.load(f, allow_pickle=True)
            assert raw.dtype == np.float64 and raw.shape == ((T, 512) if ncrops == 1 else (T, 10, 512))
            feats = C.load_features(f)                      # dataset.py:53-55
            test_item = C.getitem_test(feats)               # dataset.py:68-86
            train_item = C.getitem_train(feats)             # dataset.py:87-99
            assert test_item.shape == (T, ncrops, 513) and train_item.shape == (ncrops, 32, 513)
            assert test_item.dtype == np.float32 and train_item.dtype == np.float32
            assert np.allclose(test_item[..., 512], np.linalg.norm(feats.reshape(T, ncrops, 512), axis=2), rtol=1e-6)
            assert np.isfinite(train_item).all() and float(np.abs(train_item[..., :512]).max()) > 0
            # T < 32: every segment of process_feat is a single snippet row, in order (utils.py:39-42)
            r = np.linspace(0, T, 33, dtype=int)
            crop0 = feats.reshape(T, ncrops, 512)[:, 0]
            for s in (0, 15, 31):
                want = crop0[r[s]:r[s + 1]].mean(0) if r[s] != r[s + 1] else crop0[min(r[s], T - 1)]
                assert np.allclose(train_item[0, s, :512], want, rtol=1e-6)


def test_second_device_after_first():
    """the > 48 KB shared-memory opt-in and the SM count are per DEVICE; a process that used cuda:0 must be
    able to run on cuda:1 (needs a 2-GPU box: `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tedspad_b200.extraction import SnippetExtractor
    name = "unet_r3d18_112"
    arch, hw, reso, _, _ = _cases.CASES[name]
    rows = []
    for dev in ("cuda:0", "cuda:1"):
        fa, ft = _modules(name)
        ext = SnippetExtractor(fa.to(dev), ft.to(dev), reso=reso, batch_clips=2)
        with torch.cuda.device(dev):
            clip = _cases.case_clip(name)
            rows.append(ext.extract_video(torch.from_numpy(np.ascontiguousarray(clip)).repeat(2, 1, 1, 1)))
    assert np.array_equal(rows[0], rows[1])


def test_drop_in_reference_script_lines_unetpp_largei3d(tmp_path, monkeypatch):
    """dali_extraction.py:109-110,122-123 (model construction from checkpoints, the DEFAULT arch='unet++' + 'largei3d'
    with kin_pretrained=True) and :168-179 (loop body) run verbatim against this package: one checkpoint file holding
    both state dicts, a DataParallel-style 'module.' prefix on the anonymizer's, the Kinetics file at
    '../saved_models/i3d_r50_kinetics.pth' relative to the script's working directory."""
    from aux_code.model_loaders import load_fa_model, load_ft_model
    name = "unetpp_largei3d_224"
    sd_fa, sd_ft = _cases.case_weights(name)
    (tmp_path / "saved_models").mkdir()
    (tmp_path / "feature_extraction").mkdir()
    kin = {k[len("i3d."):]: v for k, v in sd_ft.items() if k.startswith("i3d.") and not k.startswith("i3d.fc.")}
    kin["fc.weight"], kin["fc.bias"] = torch.zeros(400, 2048), torch.zeros(400)
    torch.save(kin, tmp_path / "saved_models" / "i3d_r50_kinetics.pth")
    torch.save({"fa_model_state_dict": {"module." + k: v for k, v in sd_fa.items()}, "ft_model_state_dict": sd_ft, "epoch": 20},
               tmp_path / "saved_models" / "model_20_bestAcc_0.7504.pth")
    monkeypatch.chdir(tmp_path / "feature_extraction")
    anonymized = True
    saved_fa_model = os.path.join('..', 'saved_models', 'model_20_bestAcc_0.7504.pth') if anonymized else None
    saved_ft_model = os.path.join('..', 'saved_models', 'model_20_bestAcc_0.7504.pth')
    fa_model = load_fa_model(arch='unet++', saved_model_file=saved_fa_model)
    ft_model = load_ft_model(arch='largei3d', kin_pretrained=True, saved_model_file=saved_ft_model, num_classes=102)
    ft_model.to(device=torch.device(0))
    fa_model.to(device=torch.device(0))
    ft_model.eval()
    fa_model.eval()
    clip = _cases.case_clip(name)
    x_ref, enc_ref, f_ref = _cases.oracle_features(name, clip)
    inputs = x_ref.unsqueeze(0).cuda()
    with torch.no_grad():
        ori_bs, ori_t, ori_c, ori_h, ori_w = inputs.permute(0, 2, 1, 3, 4).shape
        inputs = inputs.view(-1, inputs.shape[2], inputs.shape[3], inputs.shape[4])
        inputs = fa_model(inputs).reshape(ori_bs, ori_t, ori_c, ori_h, ori_w)
        try:
            output = ft_model.extract_features(inputs)
        except:  # noqa: E722
            output = ft_model.i3d.extract_features(inputs)
        row = output.squeeze().cpu().numpy()
    e = (inputs.cpu() - enc_ref).abs()
    m = _cases.parity_metrics(torch.from_numpy(row), f_ref)
    print(f"\nunet++ drop-in: anonymized err max={e.max():.4f}; features cos={m['cos']:.6f} max_abs={m['max_abs']:.4f}")
    assert row.shape == (2048,) and e.max() < 0.1
    assert m["cos"] >= COS_GATE and m["max_abs"] <= ABS_GATE, m


@pytest.mark.parametrize("name", list(_cases.CONSUMER_CASES))
def test_device_side_mgfn_consumer_matches_reference_getitem(name):
    """tedspad_b200.consumer (process_feat + magnitude + [ncrops,T,F] layout on the device) against the
    golden outputs of the reference's own Dataset.__getitem__ (tests/golden/consumer_v1.npz), fp32 tolerance 1e-6."""
    from tedspad_b200 import consumer
    G2 = np.load(_cases.CONSUMER_GOLDEN)
    feats = torch.from_numpy(_cases.consumer_case_features(name)).cuda()      # float64, as the .npy files hold them
    test_item = consumer.getitem_test(feats).cpu().numpy()
    train_item = consumer.getitem_train(feats).cpu().numpy()
    for got, want in ((test_item, G2[f"{name}/test"]), (train_item, G2[f"{name}/train"])):
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.allclose(got, want, rtol=2e-6, atol=1e-6), float(np.abs(got - want).max())
    with pytest.raises(RuntimeError, match="CUDA"):
        consumer.getitem_test(torch.zeros(4, 8))


@pytest.mark.parametrize("arch,hw", [("unet", (104, 136)), ("unet++", (96, 144))])
def test_anonymizer_only_streaming_at_native_resolution(arch, hw):
    """SURVEY 8f-4 (visualization/visualize_anonymization.py:65-115): every frame of a video through the anonymizer at
    its native resolution - sizes that are neither 112 nor 224, for the UNet not even a multiple of 16 (the F.pad branch
    of unet_parts.py:57-63) - in chunks, against the fp32 oracle; then the script's colour flip and min-max uint8."""
    from aux_code.model_loaders import load_fa_model
    from tedspad_b200 import visualization as V
    H, W = hw
    frames = torch.from_numpy(M.structured_clip_u8(31, 10, H, W))                   # uint8 [10,H,W,3]
    x = frames.permute(0, 3, 1, 2).float() / 255.0                                   # ToPILImage -> ToTensor
    with torch.no_grad():
        sd = M.calibrated_state_dict(arch, 21, x)
        ref = M.anonymizer_forward(arch, sd, x)
    fa = load_fa_model(arch=arch)
    fa.load_state_dict(sd, strict=True)
    fa = fa.cuda().eval()
    got = V.anonymize_frames(fa, frames, chunk=4)
    assert got.shape == (10, 3, H, W) and got.dtype == torch.float32 and got.is_cuda
    want = torch.flip(ref, dims=[1])
    err = (got.cpu() - want).abs()
    scale = float(want.abs().max())
    print(f"\n{arch} {H}x{W}: max err {err.max():.4f}, rms {err.pow(2).mean().sqrt():.5f}, |ref|max {scale:.3f}")
    assert err.max() < 0.06 * max(scale, 1.0) and err.pow(2).mean().sqrt() < 0.012 * max(scale, 1.0)
    vid = V.to_uint8_video(got)
    ref_vid = V.to_uint8_video(want)
    assert vid.shape == (10, H, W, 3) and vid.dtype == np.uint8
    assert np.abs(vid.astype(np.int32) - ref_vid.astype(np.int32)).mean() < 2.0
    # the module call the script itself makes (fa_model(inputs) on float frames) gives the same images
    with torch.no_grad():
        direct = fa(x[:3].cuda())
    assert (direct - torch.flip(got[:3], dims=[1])).abs().max() < 0.02 * max(scale, 1.0)


def test_wrapper_i3d_forward_logits_and_embedding():
    """The module protocol remnants of VERDICT r1: wrapper_i3d.forward (model_loaders.py:265-268) and I3Res50.forward
    (large_i3d.py:229-246) in eval mode - logits through the fc head, the 128-d L2-normalised mlp embedding."""
    name = "unet_largei3d_224"
    _, ft = _modules(name)
    sd_ft = _cases.case_weights(name)[1]
    clip = _cases.case_clip(name)
    _, enc_ref, _ = _cases.oracle_features(name, clip)
    with torch.no_grad():
        pred_ref, emb_ref = M.wrapper_i3d_forward(sd_ft, enc_ref)
        pred, emb = ft(enc_ref.cuda())
        logits, feat = ft.i3d(enc_ref.cuda())
    assert pred.shape == (1, 102) and emb.shape == (1, 128) and feat.shape == (2048,)
    mp, me = _cases.parity_metrics(pred.cpu(), pred_ref), _cases.parity_metrics(emb.cpu(), emb_ref)
    print(f"\nwrapper_i3d.forward: pred cos={mp['cos']:.6f} max_abs={mp['max_abs']:.4f}; embedding cos={me['cos']:.6f} "
          f"max_abs={me['max_abs']:.4f} norm={float(emb.norm()):.5f}")
    assert mp["cos"] >= 0.999 and me["cos"] >= 0.999 and abs(float(emb.norm()) - 1.0) < 1e-3
    assert torch.equal(logits, pred)
