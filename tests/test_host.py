"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
boundary modules mirror the reference's parameter tree, weight packing, driver logic, sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import _cases
import _emu
from oracle import models as M
from oracle import preprocess as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))

from tedspad_b200 import _lib as L, engine, extraction, ops  # noqa: E402
from aux_code.model_loaders import load_fa_model, load_ft_model  # noqa: E402


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tedspad.h")).read()
    declared = set(re.findall(r"\b(tedspad_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = L.lib()  # raises if the .so is missing: there is no fallback
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.tedspad_abi_version() == L.ABI_VERSION
    lay = (ctypes.c_int32 * 10)()
    assert lib.tedspad_abi_layout(lay, 10) == 10
    assert list(lay) == [ctypes.sizeof(L.TensorDesc), ctypes.sizeof(L.ConvDesc), ctypes.sizeof(L.ConvSlabDesc),
                         ctypes.sizeof(L.SlabPlan), L.ConvDesc.y2.offset, L.ConvSlabDesc.kind.offset,
                         L.ConvSlabDesc.res.offset, L.ConvSlabDesc.oc_clip.offset, L.ConvSlabDesc.stack_rows.offset,
                         L.SlabPlan.tab.offset], "the ctypes structs drifted from include/tedspad.h as compiled"
    assert ctypes.sizeof(L.TensorDesc) == 48 and ctypes.sizeof(L.ConvDesc) == 48 * 2 + 24 + 19 * 4 + 4 + 2 * 48 + 8
    assert ctypes.sizeof(L.ConvSlabDesc) == 384 and ctypes.sizeof(L.SlabPlan) == 1168


def test_library_has_blackwell_sass():
    out = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out and "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out


@pytest.mark.parametrize("arch", ["unet", "unet++", "i3d", "largei3d", "r3d_18"])
def test_state_dict_matches_reference_tree(arch):
    """keys + shapes equal the oracle's spec, which make_golden.py loaded strict=True into the real reference
    (unet++: smp is not installable here - the spec restates smp 0.3.3's published module tree; the encoder half is
    checked against torchvision's resnet18 below)."""
    mod = load_fa_model(arch=arch) if arch in ("unet", "unet++") else load_ft_model(arch=arch, num_classes=102)
    sd = mod.state_dict()
    want = {}
    for wk, bk, bn, shape in M.conv_bn_names(arch):
        want[wk] = tuple(shape)
        if bk:
            want[bk] = (shape[0],)
        if bn:
            for s in ("weight", "bias", "running_mean", "running_var"):
                want[f"{bn}.{s}"] = (shape[0],)
            want[f"{bn}.num_batches_tracked"] = ()
    for k, shape, _ in M.extra_params(arch, 102):
        want[k] = tuple(shape)
    assert set(sd) == set(want), sorted(set(sd) ^ set(want))[:6]
    for k, v in sd.items():
        assert tuple(v.shape) == want[k], (k, tuple(v.shape), want[k])
    if arch == "i3d":
        assert list(sd)[0] == "logits.conv3d.weight"  # registered before the trunk (i3d.py:298-304)
    if arch == "unet++":
        # smp's ResNetEncoder IS torchvision's ResNet minus fc (avgpool has no parameters): same keys, same shapes
        import torchvision
        tv = {k: tuple(v.shape) for k, v in torchvision.models.resnet18().state_dict().items() if not k.startswith("fc.")}
        mine = {k[len("encoder."):]: tuple(v.shape) for k, v in sd.items() if k.startswith("encoder.")}
        assert mine == tv
        assert [k for k in sd if not k.startswith("encoder.")][0] == "decoder.blocks.x_0_0.conv1.0.weight"
        assert tuple(sd["decoder.blocks.x_0_2.conv1.0.weight"].shape) == (64, 320, 3, 3)
        assert tuple(sd["segmentation_head.0.weight"].shape) == (3, 32, 3, 3) and "segmentation_head.0.bias" in sd


def test_checkpoint_formats(tmp_path):
    """'module.' prefix strip (model_loaders.py:43-46), FrozenBN 'scale' rename (:80), .i3d fallback (:84)."""
    fa = load_fa_model(arch="unet")
    p = tmp_path / "fa.pth"
    torch.save({"fa_model_state_dict": {"module." + k: v for k, v in fa.state_dict().items()}, "epoch": 3}, p)
    fa2 = load_fa_model(saved_model_file=str(p), arch="unet")
    assert all(torch.equal(a, b) for a, b in zip(fa.state_dict().values(), fa2.state_dict().values()))
    ft = load_ft_model(arch="largei3d", num_classes=102)
    inner = {k[len("i3d."):]: v for k, v in ft.state_dict().items() if k.startswith("i3d.")}
    p2 = tmp_path / "ft.pth"
    torch.save({"ft_model_state_dict": inner}, p2)
    ft2 = load_ft_model(arch="largei3d", saved_model_file=str(p2), num_classes=102)
    assert torch.equal(ft2.i3d.conv1.weight, ft.i3d.conv1.weight)
    assert not hasattr(ft2, "extract_features") and hasattr(ft2.i3d, "extract_features")
    # the default arch (what dali_extraction.py:122 / st_feature_extraction.py:72 request), DataParallel-saved
    fpp = load_fa_model()
    p3 = tmp_path / "fa_pp.pth"
    torch.save({"fa_model_state_dict": {"module." + k: v for k, v in fpp.state_dict().items()}}, p3)
    fpp2 = load_fa_model(saved_model_file=str(p3))
    assert type(fpp2).__name__ == "UnetPlusPlus"
    assert all(torch.equal(a, b) for a, b in zip(fpp.state_dict().values(), fpp2.state_dict().values()))


def test_no_cpu_fallback():
    fa = load_fa_model(arch="unet").eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        fa(torch.zeros(1, 3, 16, 16))
    ft = load_ft_model(arch="i3d", num_classes=102).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        ft.extract_features(torch.zeros(1, 3, 16, 224, 224))
    with pytest.raises(RuntimeError, match="CUDA"):
        load_fa_model(arch="unet++").eval()(torch.zeros(1, 3, 32, 32))


def test_packed_conv_layout():
    w = torch.arange(2 * 3 * 1 * 2 * 2, dtype=torch.float32).reshape(2, 3, 1, 2, 2) / 64
    pc = ops.PackedConv(w, None, None, cin_pad=8, device="cpu")
    assert pc.w.shape == (16, 64) and pc.k_pad == 64 and pc.n_tile == 16
    k = 0
    for a in range(2):
        for b in range(2):
            for c in range(8):
                exp = float(w[1, c, 0, a, b]) if c < 3 else 0.0
                assert float(pc.w[1, k]) == exp
                k += 1
    assert ops.n_tiling(288) == (144, 288) and ops.n_tiling(512) == (256, 512) and ops.n_tiling(102) == (112, 112)
    g = torch.tensor([2.0, 0.5]); beta = torch.tensor([0.1, -0.2]); mu = torch.tensor([1.0, 2.0]); var = torch.tensor([4.0, 0.25])
    pcb = ops.PackedConv(w, torch.tensor([0.5, 0.25]), (g, beta, mu, var, 0.0), cin_pad=8, device="cpu")
    assert torch.allclose(pcb.bias[:2], (torch.tensor([0.5, 0.25]) - mu) * g / var.sqrt() + beta)
    assert torch.allclose(pcb.w[0, 0].float(), (w[0, 0, 0, 0, 0] * 1.0).to(torch.bfloat16).float())


def _emulated(monkeypatch, fp32=False):
    """Replace the CUDA operators by tests/_emu.py; fp32=True also stores activations/weights in fp32 so
    that executor wiring can be checked exactly (no rounding) against the oracle."""
    for n in ("conv_forward", "conv_slab_forward", "planes_to_clip", "maxpool", "upsample2x", "outconv_sigmoid",
              "avgpool_features", "nchw_to_cl", "preprocess", "upsample2x_nearest", "frames_to_clip"):
        monkeypatch.setattr(ops, n, getattr(_emu, n))
    if fp32:
        orig = ops.CLTensor.__init__

        def init(self, *a, **k):
            if k.get("dtype", torch.bfloat16) == torch.bfloat16:
                k["dtype"] = torch.float32
            orig(self, *a, **k)
        monkeypatch.setattr(ops.CLTensor, "__init__", init)
        monkeypatch.setattr(ops, "BF16", torch.float32)


def test_executor_wiring_unet_r3d_emulated(monkeypatch):
    """The executors' graph wiring (slices, halos, pads, residuals) against the oracle, with the CUDA
    operators replaced by tests/_emu.py (bf16 storage, fp32 math).  Also predicts the bf16 error budget."""
    _emulated(monkeypatch)
    name = "unet_r3d18_112"
    sd_fa, sd_ft = _cases.case_weights(name)
    x, enc_ref, feat_ref = _cases.oracle_features(name, _cases.case_clip(name))
    ue = engine.UNetExecutor(sd_fa, "cpu")
    x0 = ue.input_buffer(16, 112, 112)
    ops.nchw_to_cl(x, x0)
    enc = ops.CLTensor(1, 16, 112, 112, engine.ENC_IN_CHANNELS, device="cpu")
    enc.buf.zero_()
    fr = torch.empty(16, 3, 112, 112)
    ue.run(x0, enc, 16, fr)
    err = (fr - enc_ref.reshape(16, 3, 112, 112)).abs()
    assert err.max() < 0.08 and err.pow(2).mean().sqrt() < 0.01
    pred, feat = engine.R3D18Executor(sd_ft, "cpu").run(enc)
    m = _cases.parity_metrics(feat[0], feat_ref)
    assert m["cos"] > 0.9995 and m["max_abs"] < 2e-2, m


def _cl_fp32(x):
    n, c, d, h, w = x.shape
    t = ops.CLTensor(n, d, h, w, engine.ENC_IN_CHANNELS, device="cpu")  # 4: the SLAB stem's clip layout
    t.buf.zero_()
    t.interior()[..., :c] = x.permute(0, 2, 3, 4, 1)
    return t


def test_executor_wiring_i3res50_exact(monkeypatch):
    """fp32-storage emulation: I3Res50Executor's graph == large_i3d.py:249-263 to fp32 round-off
    (odd 55x55 / 27x27 extents, (2,3,3) pad-0 pool, temporal convs, strided downsample, residuals)."""
    _emulated(monkeypatch, fp32=True)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 8, 96, 96, generator=g)
    sd = M.calibrated_state_dict("largei3d", 9, x)
    ref = M.i3res50_extract_features(sd, x).flatten()
    feat = engine.I3Res50Executor(sd, "cpu").run(_cl_fp32(x))
    m = _cases.parity_metrics(feat.flatten(), ref)
    assert m["max_abs"] < 1e-4 and m["cos"] > 0.999999, m


def test_executor_wiring_unet_exact_odd_size(monkeypatch):
    """fp32-storage emulation of the UNet at a size not divisible by 16 (F.pad branch of unet_parts.py:57-63)."""
    _emulated(monkeypatch, fp32=True)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 40, 52, generator=g)
    sd = M.calibrated_state_dict("unet", 6, x)
    ref = M.unet_forward(sd, x)
    ue = engine.UNetExecutor(sd, "cpu")
    x0 = ue.input_buffer(2, 40, 52)
    ops.nchw_to_cl(x, x0)
    enc = ops.CLTensor(2, 1, 40, 52, 8, device="cpu")
    out = torch.empty(2, 3, 40, 52)
    ue.run(x0, enc, 1, out)
    assert (out - ref).abs().max() < 1e-4


def test_executor_wiring_unetpp_exact(monkeypatch):
    """fp32-storage emulation of the UNet++ executor (shared per-resolution concat buffers, permuted conv1 input
    channels, slot E re-use, the full-resolution tail + head computed at half resolution in space-to-depth form)
    against the oracle restatement of smp's UnetPlusPlus, to fp32 round-off; non-square frame."""
    _emulated(monkeypatch, fp32=True)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 48, 80, generator=g)
    sd = M.calibrated_state_dict("unet++", 12, x)
    taps = {}
    ref = M.unetpp_forward(sd, x, taps=taps)
    ex = engine.UNetPPExecutor(sd, "cpu")
    x0 = ex.input_buffer(4, 48, 80)
    ops.nchw_to_cl(x, x0)
    enc = ops.CLTensor(2, 2, 48, 80, engine.ENC_IN_CHANNELS, device="cpu")
    out = torch.empty(4, 3, 48, 80)
    ex.run(x0, enc, 2, out)
    assert (out - ref).abs().max() < 1e-4
    for name, buf in (("decoder.blocks.x_1_2.conv2", ex.bufs.find("P2").slice(128, 64)),
                      ("decoder.blocks.x_0_1.conv2", ex.bufs.find("x_0_1")), ("encoder.layer3.1", ex.bufs.find("f16"))):
        assert (buf.to_ncdhw()[:, :, 0] - taps[name]).abs().max() < 1e-4, name
    # glue: the clip is the RAW reshape of the frames (dali_extraction.py:171-173)
    want = ref.reshape(2, 2, 3, 48, 80).reshape(2, 3, 2, 48, 80)
    assert (enc.to_ncdhw()[:, :3] - want).abs().max() < 1e-4
    # without frames_out the head's KX epilogue scatters straight into the clip (no head tensor, no glue kernel): same clip
    assert ops.slab_runs_kx(ex.bufs.find("x_0_3"), ex.head.slab)
    enc2 = ops.CLTensor(2, 2, 48, 80, engine.ENC_IN_CHANNELS, device="cpu")
    enc2.buf.zero_()
    ex.run(x0, enc2, 2)
    assert torch.equal(enc2.interior()[..., :3], enc.interior()[..., :3]) and bool((enc2.interior()[..., 3:] == 0).all())
    with pytest.raises(RuntimeError, match="divisible by 16"):
        ex.run(ex.input_buffer(1, 40, 52), ops.CLTensor(1, 1, 40, 52, engine.ENC_IN_CHANNELS, device="cpu"), 1)


def test_executor_wiring_i3d_trunk_exact(monkeypatch):
    """fp32-storage emulation of the Inception trunk on a small odd-sized clip (TF-SAME asymmetric pads,
    branch concat by channel slices) against the oracle's Mixed_5c tap."""
    _emulated(monkeypatch, fp32=True)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(1, 3, 16, 72, 88, generator=g)
    sd = M.calibrated_state_dict("i3d", 8, torch.rand(1, 3, 16, 224, 224, generator=g), feature_mean=None)
    taps = {}
    with torch.no_grad():
        try:
            M.i3d_extract_features(sd, x, taps=taps)
        except RuntimeError:
            pass  # AvgPool3d([2,7,7]) cannot run on the small map; the trunk taps are what we compare
    ex = engine.I3DExecutor(sd, "cpu")
    fmap = ex.run_trunk(_cl_fp32(x))
    m = _cases.parity_metrics(fmap.to_ncdhw(), taps["Mixed_5c"])
    assert m["max_abs"] < 1e-3 and m["cos"] > 0.999999, m
    with pytest.raises(RuntimeError, match="AvgPool3d"):
        ex.run(_cl_fp32(x))


def test_snippet_indexing_and_crops_match_oracle():
    for n in (0, 1, 10, 16, 31, 32, 33, 64, 70, 100, 7001):
        d = extraction.dali_snippet_frames(n)
        assert d.tolist() == M.dali_snippet_frames(n) or (n == 0 and d.shape[0] == 0)
        s = extraction.shanghai_snippet_frames(n)
        assert s.tolist() == np.asarray(M.shanghai_snippet_frames(n)).reshape(-1, 16).tolist()
    for (h, w) in ((240, 320), (480, 856), (360, 640)):
        for nc in (1, 5, 10):
            (ch, cw), boxes = extraction.crop_boxes(h, w, nc)
            assert (ch, cw) == P.crop_size(h, w) and boxes == P.multi_crop_boxes(h, w, ch, cw, nc)
    assert extraction.crop_boxes(480, 856, 1, square_from_h=True)[0] == (384, 384)
    assert extraction.crop_boxes(480, 856, 1, square_from_h=True)[1] == [(48, 236, 0)]
    # no_ar_distortion: DALI path crops the square of the short side (dali_extraction.py:46); the ShanghaiTech path
    # takes min over the (H, W, 3) array shape = 3 (shanghai_dl.py:30) - reproduced as written, like the oracle
    assert extraction.crop_boxes(240, 320, 1, no_ar_distortion=True)[0] == P.crop_size(240, 320, no_ar_distortion=True) == (192, 192)
    assert extraction.crop_boxes(480, 856, 1, no_ar_distortion=True, square_from_h=True)[0] == (2, 2)
    f = np.random.RandomState(0).randint(0, 256, (480, 856, 3)).astype(np.uint8)
    assert P.shanghai_augmentation(f, no_ar_distortion=True).shape == (3, 224, 224)


def test_feature_path_and_resume(tmp_path):
    assert extraction.feature_path("out", "/d/Videos/Abuse/Abuse001_x264.mp4") == os.path.join("out", "Abuse001_x264.npy")
    assert extraction.feature_path("out", "/d/01_001.avi") == os.path.join("out", "01_001.npy")

    class Fake:
        calls = []

        def extract_video(self, frames):
            Fake.calls.append(int(frames.shape[0]))
            return np.full((extraction.dali_snippet_frames(frames.shape[0]).shape[0], 4), float(frames.shape[0]))

    vids = [(f"/x/v{i}.mp4", 40 + 10 * i, (lambda n=40 + 10 * i: torch.zeros(n, 2, 2, 3, dtype=torch.uint8))) for i in range(5)]
    np.save(tmp_path / "v2.npy", np.zeros((1, 4)))  # already extracted -> skipped (dali_extraction.py:121)
    written = extraction.extract_dataset(Fake(), vids, str(tmp_path), log=lambda *_: None)
    assert sorted(os.path.basename(w) for w in written) == ["v0.npy", "v1.npy", "v3.npy", "v4.npy"]
    assert 60 not in Fake.calls
    a = np.load(tmp_path / "v4.npy")
    assert a.dtype == np.float64 and a.shape == (3, 4)


def test_shard_videos_properties():
    rs = np.random.RandomState(0)
    lengths = rs.randint(100, 30000, 97).tolist()
    for ws in (1, 2, 4, 8):
        shards = extraction.shard_videos(lengths, ws)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(97))                      # a partition: every video exactly once
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)      # LPT balance bound


# ------------------------------------------------------------------------- N > 1 host logic (gloo, CPU)
class _StubExtractor:
    """Deterministic stand-in for SnippetExtractor.extract_video (the CUDA path cannot run here): features are a
    function of the frames only, so sharded and single-process runs must produce identical files."""

    def extract_video(self, frames):
        n = extraction.dali_snippet_frames(frames.shape[0]).shape[0]
        base = frames.reshape(frames.shape[0], -1).double().mean(1)
        return np.stack([np.full(8, float(base[min(32 * i, frames.shape[0] - 1)]) + i) for i in range(n)]) if n else np.zeros((0, 8))


def _stub_videos():
    rs = np.random.RandomState(3)
    vids = []
    for i, n in enumerate(rs.randint(20, 400, 13).tolist()):
        def loader(i=i, n=n):
            g = torch.Generator().manual_seed(100 + i)
            return torch.randint(0, 256, (n, 4, 4, 3), generator=g, dtype=torch.uint8)
        vids.append((f"/data/vid_{i:02d}.mp4", n, loader))
    return vids


class _FailingExtractor(_StubExtractor):
    def extract_video(self, frames):
        import torch.distributed as dist
        if dist.get_rank() == 1:
            raise ValueError("decoder exploded")
        return super().extract_video(frames)


def _gloo_failing_worker(rank, world, port, folder, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            extraction.extract_dataset_distributed(_FailingExtractor(), _stub_videos(), folder, log=lambda *_: None)
            q.put((rank, "no error"))
        except RuntimeError as e:
            q.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_failure_is_reported_on_every_rank(tmp_path):
    """A rank whose shard fails must not leave the others waiting in the manifest gather: the error travels with the
    gather and every rank raises (bench.py relies on this to keep its JSON line when an extra fails)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_failing_worker, args=(r, 2, port, str(tmp_path / "out"), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert set(got) == {0, 1}
    assert all("rank 1: ValueError: decoder exploded" in m for m in got.values()), got


def _gloo_worker(rank, world, port, folder, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        manifest = extraction.extract_dataset_distributed(_StubExtractor(), _stub_videos(), folder, log=lambda *_: None)
        q.put((rank, manifest))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharded_extraction_equals_single(tmp_path):
    """Two processes over gloo: disjoint shards, every file written exactly once, no data-path collective, the
    gathered manifest identical on both ranks, and the files bit-identical to a single-process run."""
    import socket
    import torch.multiprocessing as mp
    single = tmp_path / "single"
    extraction.extract_dataset(_StubExtractor(), _stub_videos(), str(single), log=lambda *_: None)
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    sharded = tmp_path / "sharded"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, str(sharded), q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert results[0] == results[1] and len(results[0]) == 13
    expect = extraction.shard_videos([v[1] for v in _stub_videos()], 2)
    for r in (0, 1):
        mine = sorted(os.path.basename(f) for f, owner in results[0].items() if owner == r)
        assert mine == sorted(f"vid_{i:02d}.npy" for i in expect[r])
    for f in sorted(os.listdir(single)):
        assert np.array_equal(np.load(single / f), np.load(sharded / f)), f
    assert sorted(os.listdir(single)) == sorted(os.listdir(sharded))


def test_segment_features_match_reference_statements():
    """a-11: the (disabled) 32-segment pooling of both scripts, against their statement-for-statement restatement."""
    import warnings
    rs = np.random.RandomState(5)
    for n in (1, 7, 31, 32, 33, 64, 219, 1000):
        f = rs.rand(n, 16) + 0.1
        assert np.array_equal(extraction.segment_features(f, 16), M.segment_features_dali(f, 16))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = M.segment_features_shanghai(f)
        got = extraction.segment_features(f, rule="shanghai")
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(want)), n
    with pytest.raises(IndexError):
        extraction.segment_features(np.zeros((0, 16)))


def test_extract_videos_packing_bookkeeping_on_cpu():
    """The cross-video batch packer of SnippetExtractor.extract_videos (frame-range offsets inside the shared staging
    buffer, per-batch metadata, completion order, empty videos, flush on a frame-size change) with the CUDA pipeline
    replaced by a deterministic function of exactly the frames each clip references."""
    class FakePipe(extraction.SnippetExtractor):
        def __init__(self, ncrops, batch_clips, source="dali"):   # no modules, no device: only the driver logic
            self.reso, self.T, self.skip = (8, 8), 16, 2
            self.cf, self.no_ar, self.ncrops, self.source, self.batch_clips = 0.8, False, ncrops, source, batch_clips
            self.batches_seen = []

        def features_stream(self, batches):
            for chunks, desc, crop_hw in batches:
                frames = torch.cat([c for c in (chunks if isinstance(chunks, (list, tuple)) else [chunks])], 0).double()
                self.batches_seen.append(desc.shape[0] // self.T)
                d = desc.reshape(-1, self.T, 4)
                rows = []
                for clip in d:
                    acc = [0.0, 0.0, 0.0]
                    for t, (fi, top, left, flip) in enumerate(clip.tolist()):
                        if fi >= 0:
                            acc[0] += float(frames[fi].mean()) * (t + 1)
                            acc[1] += float(frames[fi, top % frames.shape[1], left % frames.shape[2], 0]) + flip
                    acc[2] = float(crop_hw[0] * 1000 + crop_hw[1])
                    rows.append(acc)
                yield torch.tensor(rows, dtype=torch.float32).reshape(len(rows), 1, 3)

    rs = np.random.RandomState(0)
    def vid(n, h, w):
        return torch.from_numpy(rs.randint(0, 256, (n, h, w, 3)).astype(np.uint8))
    vids = [vid(40, 20, 24), vid(130, 20, 24), vid(0, 20, 24), vid(33, 20, 24), vid(70, 16, 30), vid(95, 16, 30), vid(31, 20, 24)]
    for ncrops, bc, source in ((1, 4, "dali"), (10, 20, "dali"), (5, 7, "dali"), (1, 3, "shanghai")):
        ext = FakePipe(ncrops, bc, source)
        packed = list(ext.extract_videos(iter(vids)))
        assert [i for i, _ in packed] == list(range(len(vids)))
        n_batches_packed = len(ext.batches_seen)
        per_video = 0
        for (i, got), v in zip(packed, vids):
            if ext.snippet_frames(v.shape[0]).shape[0] == 0:
                assert got.shape[0] == 0
                continue
            ext.batches_seen = []
            want = ext.extract_video(v)
            per_video += len(ext.batches_seen)
            assert got.shape == want.shape and np.array_equal(got, want), (ncrops, bc, source, i)
        assert n_batches_packed <= per_video      # packing never issues more batches than the per-video path


def test_needed_frames_staging_property():
    """SnippetExtractor._needed_frames, both branches (contiguous range for small frames; only the frames a snippet
    keeps, one chunk per run of consecutive frames, for large ones - the ShanghaiTech reader keeps every second frame,
    shanghai_dl.py:73): the staged chunks indexed by the returned positions are exactly the frames the snippets name,
    zero-image slots (-1, the DALI tail padding) stay -1, and the staged-frame count is what the chunks hold."""
    class Bare(extraction.SnippetExtractor):
        def __init__(self):
            pass
    ext = Bare()
    rs = np.random.RandomState(3)
    frames = torch.from_numpy(rs.randint(0, 256, (300, 6, 5, 3)).astype(np.uint8))
    cases = 0
    for sparse_bytes in (10 ** 9, 1):           # never sparse / always sparse (when the index set has gaps)
        ext.SPARSE_FRAME_BYTES = sparse_bytes
        for source, n in (("dali", 300), ("dali", 47), ("shanghai", 300), ("shanghai", 100), ("dali", 16)):
            snips = (extraction.dali_snippet_frames(n) if source == "dali" else extraction.shanghai_snippet_frames(n))
            for s0 in range(0, snips.shape[0], 3):
                sn = snips[s0:s0 + 3]
                for base in (0, 17):
                    chunks, rel, n_staged = ext._needed_frames(frames[:n], sn, base)
                    staged = torch.cat(list(chunks), 0)
                    assert staged.shape[0] == n_staged
                    assert np.array_equal(rel == -1, sn == -1)
                    ok = sn >= 0
                    assert rel[ok].min() >= base and rel[ok].max() < base + n_staged
                    assert torch.equal(staged[torch.from_numpy(rel[ok] - base)], frames[torch.from_numpy(sn[ok])])
                    if sparse_bytes == 1 and source == "shanghai":
                        assert n_staged == int(ok.sum())            # every second frame only: nothing unused is staged
                    cases += 1
    # an all-padding group (cannot come from the readers, but the packer must survive it)
    chunks, rel, n_staged = ext._needed_frames(frames, np.full((1, 16), -1, dtype=np.int64), 5)
    assert n_staged == 1 and (rel == -1).all() and chunks[0].shape[0] == 1
    assert cases >= 40


def test_extractor_from_reference_params_module():
    """SnippetExtractor.from_params reads the reference's own parameter file (feature_extraction/params_feature_ex.py
    restated: the GPU box has no /root/reference) and lets keyword overrides win."""
    import types
    params = types.SimpleNamespace(num_classes=102, num_frames=16, fix_skip=2, batch_size=1, reso_h=224, reso_w=224,
                                   cropping_factor=0.8, no_ar_distortion=False, num_workers=4)   # params_feature_ex.py:1-9
    ref_params = "/root/reference/feature_extraction/params_feature_ex.py"
    if os.path.exists(ref_params):   # in the build container: the restatement above IS the reference's file
        ns = {}
        exec(open(ref_params).read(), ns)
        assert {k: ns[k] for k in vars(params)} == vars(params)

    class Capture(extraction.SnippetExtractor):
        def __init__(self, fa, ft, **kw):
            self.args = (fa, ft, kw)

    ext = Capture.from_params("fa", "ft", params, ncrops=10, batch_clips=40, source="shanghai")
    assert ext.args == ("fa", "ft", dict(reso=(224, 224), num_frames=16, fix_skip=2, cropping_factor=0.8,
                                         no_ar_distortion=False, ncrops=10, batch_clips=40, source="shanghai"))
    assert Capture.from_params("fa", "ft", params, fix_skip=1).args[2]["fix_skip"] == 1


def test_activation_buffers_are_bounded_by_the_largest_batch():
    """every distinct batch size used to get its own full buffer set (OOM over a dataset's video tails).
    One allocation per name now; smaller batches are N-prefix views of it, a larger batch replaces it."""
    from tedspad_b200.engine import _Buffers
    b = _Buffers("cpu")
    full = b.get("t0", 8, 1, 6, 6, 64, (0, 1, 1))
    size0 = b.nbytes()
    for n in (8, 3, 5, 1, 7, 8, 2):
        v = b.get("t0", n, 1, 6, 6, 64, (0, 1, 1))
        assert v.N == n and v.buf.data_ptr() == full.buf.data_ptr() and v.buf.is_contiguous()
        assert tuple(v.buf.shape) == (n, 1, 8, 8, 64)
        assert b.get("t0", n, 1, 6, 6, 64, (0, 1, 1)) is v            # views are cached, not rebuilt per step
        assert b.nbytes() == size0
    assert b.find("t0") is full and b.find("t0", 3).N == 3
    big = b.get("t0", 12, 1, 6, 6, 64, (0, 1, 1))                      # larger batch: replaces, does not add
    assert big.N == 12 and b.nbytes() == size0 * 12 // 8 and len(b.pool) == 1
    other = b.get("t0", 4, 1, 10, 10, 64, (0, 1, 1))                   # new geometry under the same name: replaces
    assert (other.H, other.N) == (10, 4) and len(b.pool) == 1
    z = b.get("pad", 2, 1, 4, 4, 64, zero=True)
    assert float(z.buf.float().abs().max()) == 0.0


def test_data_parallel_is_rejected_explicitly():
    """dali_extraction.py:126-141 wraps the models in nn.DataParallel when it sees several GPUs; replicas have no
    parameters and run on threads.  The boundary modules refuse replication with a clear message."""
    from aux_code.model_loaders import load_fa_model
    fa = load_fa_model(arch="unet")
    with pytest.raises(RuntimeError, match="one process per GPU"):
        fa._replicate_for_data_parallel()


def test_u8_over_255_refinement_is_exact():
    """csrc/ops.cu u8_over_255: q = b*r; q' = fma(fma(-q, 255, b), r, q) with r = rn(1/255) = 0x3b808081 equals the
    IEEE division float(b)/255.f for all 256 bytes (exact rational arithmetic, correctly rounded at every step)."""
    from fractions import Fraction

    def rn(fr):
        a = np.float32(float(fr))
        cands = [np.nextafter(a, np.float32(-np.inf)), a, np.nextafter(a, np.float32(np.inf))]
        return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(c.view(np.uint32)) & 1))

    r = np.uint32(0x3b808081).view(np.float32)
    assert rn(Fraction(1, 255)) == r
    for b in range(256):
        q = rn(Fraction(b) * Fraction(float(r)))
        rem = rn(-Fraction(float(q)) * 255 + b)
        q2 = rn(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q)))
        assert q2 == np.float32(b) / np.float32(255.0), b


def _write_mjpg(path, n, hw=(48, 64), bgr=(10, 90, 235)):
    import cv2
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (hw[1], hw[0]))
    for i in range(n):
        fr = np.empty((hw[0], hw[1], 3), dtype=np.uint8)
        fr[..., 0], fr[..., 1], fr[..., 2] = bgr[0] + 3 * i, bgr[1], bgr[2] - 3 * i   # B ramps up, R down with the index
        wr.write(fr)
    wr.release()


def test_cv2_ingest_decodes_every_frame_in_bgr_order(tmp_path):
    """SURVEY 8f-2 (ShanghaiTech half): tedspad_b200.ingest.decode_video_cv2 = the decoding loop of
    shanghai_dl.py:43-98 - all frames, in order, BGR untouched - and the snippet indexer fed with the container's
    frame count reproduces the reference reader's clips (golden shanghai_idx/* came from its read_video)."""
    from tedspad_b200 import ingest
    G = _cases.golden()
    for n in (10, 33, 70):
        p = str(tmp_path / f"v{n}.avi")
        _write_mjpg(p, n)
        frames, total = ingest.decode_video_cv2(p)
        assert frames.dtype == torch.uint8 and tuple(frames.shape) == (n, 48, 64, 3) and total == n
        means = frames.float().mean(dim=(1, 2))                                  # [n, 3] in B, G, R order
        ramp = torch.arange(n, dtype=torch.float32) * 3
        # decoded in order, channel 0 is Blue (no RGB conversion); MJPG's chroma quantisation moves levels by a few units
        assert float((means[:, 0] - (10 + ramp)).abs().max()) <= 6 and float((means[:, 2] - (235 - ramp)).abs().max()) <= 6
        rgb, _ = ingest.decode_video_cv2(p, rgb=True)
        assert torch.equal(rgb, frames.flip(-1))
        idx = extraction.shanghai_snippet_frames(frames.shape[0], total_frames=total)
        assert np.array_equal(idx, G[f"shanghai_idx/{n}"])
    ds = ingest.cv2_dataset([str(tmp_path / "v70.avi"), str(tmp_path / "v33.avi")])
    assert [d[1] for d in ds] == [70, 33] and tuple(ds[1][2]().shape) == (33, 48, 64, 3)
    with pytest.raises(RuntimeError):
        ingest.decode_video_cv2(str(tmp_path / "missing.avi"))
    # container count and decodable frames may disagree: the reference decides skip / repeat on the container's number
    assert extraction.shanghai_snippet_frames(40, total_frames=20).shape == (2, 16)       # skip 1 (total < 32): 40 // 16
    assert extraction.shanghai_snippet_frames(12, total_frames=10).tolist()[0][-4:] == [9, 9, 9, 9]
    with pytest.raises(RuntimeError, match="could not process"):
        extraction.shanghai_snippet_frames(8, total_frames=10)


def test_graph_cache_protocol(monkeypatch):
    """engine.GraphCache: first call with a key eager, second captured + replayed, later calls replayed; a changed
    buffer generation (reallocation, new weights) forces a fresh eager call and a re-capture; profiling hooks and
    TEDSPAD_GRAPHS=0 bypass it.  (CUDA graph objects are faked: the protocol is host logic.)"""
    log = []

    class FakeGraph:
        def replay(self):
            log.append("replay")

    class FakeCapture:
        def __init__(self, g, capture_error_mode="global"):
            assert capture_error_mode == "thread_local"

        def __enter__(self):
            log.append("capture")

        def __exit__(self, *a):
            return False

    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", FakeCapture)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(engine, "USE_GRAPHS", True)
    gc = engine.GraphCache()
    calls = []

    def fn():
        calls.append(1)
        ops.LAUNCHES += 3
        return "out"

    ops.LAUNCHES = 0
    assert gc.run("k", 0, fn) == "out" and log == [] and len(calls) == 1            # eager
    assert gc.run("k", 0, fn) == "out" and log == ["capture", "replay"] and len(calls) == 2
    assert gc.run("k", 0, fn) == "out" and log == ["capture", "replay", "replay"] and len(calls) == 2
    assert ops.LAUNCHES == 3 + 3 + 3                                                  # replays count their kernels
    gc.run("k", 1, fn)                                                                # generation changed: eager again
    assert len(calls) == 3 and log[-1] == "replay"
    gc.run("k", 1, fn)
    assert log[-2:] == ["capture", "replay"]
    monkeypatch.setattr(ops, "CONV_EVENTS", [])                                       # per-launch timing: always eager
    gc.run("k", 1, fn)
    assert len(calls) == 5
    monkeypatch.setattr(ops, "CONV_EVENTS", None)
    monkeypatch.setattr(engine, "USE_GRAPHS", False)
    gc.run("k", 1, fn)
    assert len(calls) == 6


def test_prefetched_loaders_keep_order_and_overlap():
    """ingest.prefetched: results in input order, bounded look-ahead, loaders really run concurrently, an exception
    surfaces at its own position; extract_dataset(prefetch_workers=...) writes the same files as the serial path."""
    import threading
    import time
    from tedspad_b200 import ingest
    active, peak, started = [0], [0], []
    lock = threading.Lock()

    def mk(i, fail=False):
        def fn():
            with lock:
                active[0] += 1
                peak[0] = max(peak[0], active[0])
                started.append(i)
            time.sleep(0.05)
            with lock:
                active[0] -= 1
            if fail:
                raise ValueError(f"video {i}")
            return i
        return fn

    out = []
    gen = ingest.prefetched((mk(i) for i in range(10)), workers=3, depth=4)
    for v in gen:
        out.append(v)
        assert max(started) <= v + 4            # never more than `depth` ahead of the consumer
    assert out == list(range(10)) and peak[0] >= 2
    with pytest.raises(ValueError, match="video 2"):
        list(ingest.prefetched([mk(0), mk(1), mk(2, fail=True), mk(3)], workers=2))
