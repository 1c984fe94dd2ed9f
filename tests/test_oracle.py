"""CPU tests that pin the ORACLE (oracle/) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and against torchvision / Pillow themselves."""
import numpy as np
import pytest
import torch
import torchvision.transforms.functional as TF

import _cases
from oracle import models as M
from oracle import preprocess as P

G = _cases.golden()
GOLDEN_CASES = [n for n in _cases.CASES if n not in _cases.FA_ARCH]   # cases the reference itself could run here


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_features(name):
    """weights + clip regenerated from seeds, oracle forward == features of the real reference run."""
    clip = _cases.case_clip(name)
    assert str(G[f"{name}/clip_sha"]) == __import__("hashlib").sha256(np.ascontiguousarray(clip).tobytes()).hexdigest()[:16]
    x, enc_in, feat = _cases.oracle_features(name, clip)
    ref = torch.from_numpy(G[f"{name}/features"])
    m = _cases.parity_metrics(feat, ref)
    assert m["max_abs"] < 2e-4 and m["cos"] > 0.999999, m
    a = enc_in.reshape(-1).numpy()
    st = G[f"{name}/anon_stats"]
    assert abs(a.mean() - st[0]) < 1e-5 and abs(a.std() - st[1]) < 1e-5
    idx = np.random.RandomState(7).randint(0, a.size, 256)
    # anon_samples were taken from the [16,3,h,w] frames tensor, which is the same memory order as enc_in
    assert np.abs(a[idx] - G[f"{name}/anon_samples"]).max() < 1e-4


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_is_discriminative(name):
    """different clips give measurably different features (stock init would give cos 1.0)."""
    ctrl = float(G[f"{name}/control_cos"])
    assert ctrl < 0.9995, ctrl
    f1, f2 = G[f"{name}/features"], G[f"{name}/control_features"]
    assert np.linalg.norm(f1 - f2) / np.linalg.norm(f1) > 0.03


def test_glue_plane_map():
    """dali_extraction.py:171-173: encoder (channel c', time t') holds anonymizer plane 16*c' + t' = 3*t + c."""
    probe = G["glue/probe"]  # [3,16] of 10*t + c from the reference's own view/reshape
    pm = M.plane_map(16, 3)
    for (t, c), (ce, te) in pm.items():
        assert probe[ce, te] == 10 * t + c
    x = torch.arange(16 * 3 * 2 * 2, dtype=torch.float32).reshape(1, 16, 3, 2, 2)
    ident = {k: v.clone() for k, v in {}.items()}
    out = x.reshape(-1, 3, 2, 2).reshape(1, 3, 16, 2, 2)
    for (t, c), (ce, te) in pm.items():
        assert torch.equal(out[0, ce, te], x[0, t, c])


@pytest.mark.parametrize("n", [10, 16, 20, 31, 32, 33, 64, 70, 100])
def test_shanghai_indexing_matches_reference_reader(n):
    got = np.asarray(M.shanghai_snippet_frames(n), dtype=np.int64).reshape(-1, 16)
    assert np.array_equal(got, G[f"shanghai_idx/{n}"])


def test_dali_indexing():
    s = M.dali_snippet_frames(70)
    assert len(s) == 3 and s[0] == [2 * j for j in range(16)] and s[1][0] == 32
    assert s[2] == [64, 66, 68] + [-1] * 13          # pad_sequences=True: tail kept, missing frames zero
    assert len(M.dali_snippet_frames(64)) == 2 and len(M.dali_snippet_frames(65)) == 3
    assert M.dali_snippet_frames(0) == []


def test_preprocess_aa_matches_torchvision():
    rs = np.random.RandomState(0)
    for (h, w) in [(240, 320), (360, 640), (120, 160)]:
        v = rs.randint(0, 256, (2, h, w, 3)).astype(np.uint8)
        mine = P.dali_val_augmentations(v)
        vt = torch.from_numpy(v).float().permute(0, 3, 1, 2) / 255.
        ch, cw = int(h * 0.8), int(w * 0.8)
        ref = TF.resize(TF.center_crop(vt, (ch, cw)), (224, 224), antialias=True).numpy()
        assert np.abs(mine - ref).max() < 2e-6


def test_preprocess_pil_bit_exact():
    rs = np.random.RandomState(1)
    for (h, w) in [(480, 856), (300, 400), (240, 320)]:
        f = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        mine = P.shanghai_augmentation(f)
        im = TF.to_pil_image(f)
        c = int(h * 0.8)
        im = TF.resize(TF.center_crop(im, (c, c)), (224, 224), antialias=True)
        assert np.array_equal(mine, TF.to_tensor(im).numpy())


def test_multi_crop_order_matches_torchvision():
    h, w, ch, cw = 24, 32, 19, 25
    img = torch.arange(h * w, dtype=torch.float32).reshape(1, h, w)
    for n, crops in ((5, TF.five_crop(img, (ch, cw))), (10, TF.ten_crop(img, (ch, cw)))):
        boxes = P.multi_crop_boxes(h, w, ch, cw, n)
        for (t, l, fl), ref in zip(boxes, crops):
            src = img.flip(-1) if fl else img
            assert torch.equal(src[:, t:t + ch, l:l + cw], ref)
    assert P.multi_crop_boxes(h, w, ch, cw, 1)[0] == P.multi_crop_boxes(h, w, ch, cw, 5)[4]


@pytest.mark.parametrize("name", list(_cases.CONSUMER_CASES))
def test_consumer_oracle_matches_reference_getitem(name, tmp_path):
    """oracle/consumer.py == the reference's Dataset.__getitem__ (dataset.py:51-132) on the same files, bit for bit
    (golden vectors from tests/golden/make_golden_consumer.py)."""
    from oracle import consumer as C
    G2 = np.load(_cases.CONSUMER_GOLDEN)
    path = str(tmp_path / (name + ".npy"))
    np.save(path, _cases.consumer_case_features(name))
    feats = C.load_features(path)
    assert feats.dtype == np.float32
    assert np.array_equal(C.getitem_test(feats), G2[f"{name}/test"])
    assert np.array_equal(C.getitem_train(feats), G2[f"{name}/train"])


def test_crop_augmentations_match_torchvision_ten_crop():
    """oracle dali_crop_augmentations (explicit box, optional h-flip) == torchvision ten_crop + antialiased resize."""
    rs = np.random.RandomState(3)
    h, w = 120, 160
    v = rs.randint(0, 256, (2, h, w, 3)).astype(np.uint8)
    ch, cw = P.crop_size(h, w)
    vt = torch.from_numpy(v).float().permute(0, 3, 1, 2) / 255.
    crops = TF.ten_crop(vt, (ch, cw))
    for box, ref in zip(P.multi_crop_boxes(h, w, ch, cw, 10), crops):
        mine = P.dali_crop_augmentations(v, box, (ch, cw), (112, 112))
        assert np.abs(mine - TF.resize(ref, (112, 112), antialias=True).numpy()).max() < 2e-6


def test_unetpp_oracle_encoder_matches_torchvision_resnet18_and_is_discriminative():
    """arch='unet++' (smp 0.3.3, not installable: decoder parity unpinned).  What CAN be pinned here: the encoder half
    of the restatement against torchvision's resnet18 with the same weights, the output contract ([N,3,H,W], no
    activation), the divisible-by-16 check, and that the synthetic case is input-dependent."""
    import torchvision
    name = "unetpp_largei3d_224"
    sd_fa, sd_ft = _cases.case_weights(name)
    x = torch.from_numpy(P.dali_val_augmentations(_cases.case_clip(name)[:2], (224, 224)))
    r = torchvision.models.resnet18()
    r.load_state_dict({k[len("encoder."):]: v for k, v in sd_fa.items() if k.startswith("encoder.")}, strict=False)
    r.eval()
    with torch.no_grad():
        a = r.relu(r.bn1(r.conv1(x)))
        b = r.layer1(r.maxpool(a))
        c = r.layer2(b)
        d = r.layer3(c)
        feats = M.resnet18_encoder_features(sd_fa, x)
        out = M.unetpp_forward(sd_fa, x)
    for want, got in zip((a, b, c, d), feats[1:]):
        assert want.shape == got.shape and (want - got).abs().max() < 1e-4
    assert out.shape == (2, 3, 224, 224) and float(out.min()) < 0.2 and float(out.max()) > 0.8   # unbounded, image-like scale
    with pytest.raises(RuntimeError, match="divisible by 16"):
        M.unetpp_forward(sd_fa, torch.zeros(1, 3, 40, 52))
    f1 = _cases.oracle_features(name, _cases.case_clip(name))[2]
    f2 = _cases.oracle_features(name, _cases.case_clip(name, "control"))[2]
    ctrl = float(torch.nn.functional.cosine_similarity(f1, f2, dim=0))
    assert ctrl < 0.9995, ctrl


def test_wrapper_i3d_forward_oracle_matches_reference_golden():
    """oracle wrapper_i3d_forward == the unmodified reference's wrapper_i3d.forward / I3Res50.forward
    (tests/golden/wrapper_i3d_v1.npz from make_golden.py), weights and clip regenerated from seeds."""
    import os
    G3 = np.load(os.path.join(os.path.dirname(_cases.GOLDEN), "wrapper_i3d_v1.npz"))
    x = torch.rand(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(17))
    sd = M.calibrated_state_dict("largei3d", 3, x)
    with torch.no_grad():
        pred, emb = M.wrapper_i3d_forward(sd, x)
    assert np.abs(pred.numpy() - G3["pred"]).max() < 1e-4 and np.abs(emb.numpy() - G3["emb"]).max() < 1e-5
    assert np.allclose(np.linalg.norm(G3["emb"], axis=1), 1.0, atol=1e-5)
