"""Generates tests/golden/golden_v1.npz by running the UNMODIFIED reference modules imported from
/root/reference (this container only; the GPU box never sees /root/reference).

For every case it
  1. builds the seeded calibrated weights with oracle.models.calibrated_state_dict,
  2. loads them `strict=True` into the reference modules built by the reference's own
     aux_code/model_loaders.py (load_fa_model / load_ft_model) - which pins key names and shapes,
  3. runs the reference forward exactly as feature_extraction/dali_extraction.py:168-179 does
     (torchvision val_augmentations, view/reshape glue, extract_features / .i3d.extract_features),
  4. checks the oracle restatement against it (printed), and
  5. stores the reference's outputs as the golden vectors.

Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch
import torchvision.transforms.functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models as M  # noqa: E402
from oracle import preprocess as P  # noqa: E402

REF = "/root/reference"


def import_reference():
    # segmentation_models_pytorch (UnetPlusPlus) is not installed; arch='unet' never touches it.
    stub = types.ModuleType("segmentation_models_pytorch")
    stub.UnetPlusPlus = object
    sys.modules.setdefault("segmentation_models_pytorch", stub)
    sys.path.insert(0, REF)
    from aux_code.model_loaders import load_fa_model, load_ft_model
    return load_fa_model, load_ft_model


def ref_val_augmentations(video_fhwc_u8, reso=(224, 224), cropping_factor=0.8):
    """dali_extraction.py:38-50 with the same torchvision calls (DALI hands over float 0..255 FHWC)."""
    video = torch.from_numpy(video_fhwc_u8).float().unsqueeze(0)
    video = torch.transpose(video, 2, 4)
    video = torch.transpose(video, 3, 4)
    video = video / 255.
    h, w = int(video.shape[-2]), int(video.shape[-1])
    video = TF.center_crop(video.squeeze(), (int(h * cropping_factor), int(w * cropping_factor)))
    video = TF.resize(video, reso, antialias=True)
    return video.unsqueeze(dim=0)


def ref_extract(fa_model, ft_model, inputs, arch):
    """dali_extraction.py:168-179 / BASELINE config 1 for r3d_18."""
    with torch.no_grad():
        ori_bs, ori_t, ori_c, ori_h, ori_w = inputs.permute(0, 2, 1, 3, 4).shape
        frames = inputs.view(-1, inputs.shape[2], inputs.shape[3], inputs.shape[4])
        anon = fa_model(frames)
        enc_in = anon.reshape(ori_bs, ori_t, ori_c, ori_h, ori_w)
        if arch == "r3d_18":
            pred, feat = ft_model(enc_in)
            return anon, feat.squeeze(), pred.squeeze()
        try:
            output = ft_model.extract_features(enc_in)
        except AttributeError:
            output = ft_model.i3d.extract_features(enc_in)
        return anon, output.squeeze(), None


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()[:16]


sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases  # noqa: E402  (the case table, seeds and synthetic-init regimes are shared with the tests)

CASES = [(name,) + spec for name, spec in _cases.CASES.items()]


def build_case_weights(name):
    return _cases.case_weights(name)


def shanghai_index_golden():
    """Runs the reference's shanghai_frames_dataset.read_video (shanghai_dl.py:43-98) on synthetic MJPG
    clips whose frames encode their own index, and records which source frames land in which clip."""
    import tempfile
    import cv2
    sys.path.insert(0, os.path.join(REF, "feature_extraction"))
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "feature_extraction"))
    try:
        import shanghai_dl
    finally:
        os.chdir(cwd)
    ds = shanghai_dl.shanghai_frames_dataset.__new__(shanghai_dl.shanghai_frames_dataset)
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for n in (10, 16, 20, 31, 32, 33, 64, 70, 100):
            path = os.path.join(td, f"v{n}.avi")
            wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25, (64, 48))
            for i in range(n):
                wr.write(np.full((48, 64, 3), 2 * i + 10, dtype=np.uint8))  # grey level encodes the frame index
            wr.release()
            full_vid, full_pos, _ = ds.read_video(path)
            assert full_vid is not None
            # read_video records the 1-based counter of every frame it keeps (frame_pos, shanghai_dl.py:74,92);
            # the pixel values (grey = 2*i+10, shifted by MJPG's level rounding) must be monotone with it
            idx = [[p - 1 for p in clip_pos] for clip_pos in full_pos]
            grey = [[float(fr.mean()) * 255.0 for fr in clip] for clip in full_vid]
            for gi, ii in zip(grey, idx):
                assert all(abs((g - 10) / 2 - i) <= 1.5 for g, i in zip(gi, ii)), (gi, ii)
            res[f"shanghai_idx/{n}"] = np.asarray(idx, dtype=np.int64).reshape(-1, 16)
            # the product's cv2 ingest + the oracle's Pillow restatement == the reference reader's clips, bit for bit
            sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))
            from tedspad_b200 import ingest
            frames, total = ingest.decode_video_cv2(path, pin=False)
            for clip, ii in zip(full_vid, idx):
                mine = np.stack([P.shanghai_augmentation(frames[i].numpy()) for i in ii])
                assert np.array_equal(mine, clip.numpy()), n
            print(f"shanghai read_video n={n}: {len(idx)} clips; first {idx[0][:4] if idx else None}")
    return res


def wrapper_forward_golden(load_ft_model):
    """wrapper_i3d.forward / I3Res50.forward of the unmodified reference (model_loaders.py:265-268, large_i3d.py:229-246)
    on a small seeded clip -> tests/golden/wrapper_i3d_v1.npz (the oracle restatement is asserted against it here)."""
    x = torch.rand(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(17))
    sd = M.calibrated_state_dict("largei3d", 3, x)
    ft = load_ft_model(arch="largei3d", num_classes=102)
    ft.load_state_dict(sd, strict=True)
    ft.eval()
    with torch.no_grad():
        pred, emb = ft(x)
        logits, feat = ft.i3d(x)
        pred_o, emb_o = M.wrapper_i3d_forward(sd, x)
    assert (pred - pred_o).abs().max() < 1e-4 and (emb.float() - emb_o).abs().max() < 1e-5 and torch.equal(logits, pred)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wrapper_i3d_v1.npz"), pred=pred.numpy(), emb=emb.float().numpy(),
                        feat=feat.numpy())
    print("wrapper_i3d.forward: oracle == reference; wrote wrapper_i3d_v1.npz")


def main():
    load_fa_model, load_ft_model = import_reference()
    wrapper_forward_golden(load_ft_model)
    torch.set_num_threads(os.cpu_count())
    out = {}
    for name, arch, hw, reso, wseeds, cseeds in CASES:
        sd_fa, sd_ft = build_case_weights(name)
        fa = load_fa_model(arch="unet")
        ft = load_ft_model(arch=arch, num_classes=102, kin_pretrained=False)
        missing = set(fa.state_dict().keys()) ^ set(sd_fa.keys())
        assert not missing, f"fa key mismatch: {sorted(missing)[:5]}"
        missing = set(ft.state_dict().keys()) ^ set(sd_ft.keys())
        assert not missing, f"ft key mismatch: {sorted(missing)[:5]}"
        fa.load_state_dict(sd_fa, strict=True)
        ft.load_state_dict(sd_ft, strict=True)
        fa.eval(); ft.eval()
        feats = {}
        for tag, cs in (("test", cseeds[1]), ("control", cseeds[2])):
            clip = M.structured_clip_u8(cs, 16, hw[0], hw[1])
            inputs = ref_val_augmentations(clip, reso)                      # reference path
            mine = torch.from_numpy(P.dali_val_augmentations(clip, reso))  # oracle path
            d_pre = (inputs[0] - mine).abs().max().item()
            anon, feat, pred = ref_extract(fa, ft, inputs, arch)
            with torch.no_grad():
                o_enc_in = M.anonymize_and_reshape(sd_fa, mine.unsqueeze(0))
                if arch == "r3d_18":
                    o_pred, o_feat = M.r3d18_forward(sd_ft, o_enc_in)
                    o_feat = o_feat.squeeze()
                else:
                    o_feat = M.encoder_features(arch, sd_ft, o_enc_in).squeeze()
            d_anon = (anon.reshape(o_enc_in.shape) - o_enc_in).abs().max().item()
            d_feat = (feat - o_feat).abs().max().item()
            cos = torch.nn.functional.cosine_similarity(feat, o_feat, dim=0).item()
            print(f"{name}/{tag}: oracle vs reference: preprocess {d_pre:.2e}  anonymized {d_anon:.2e}  "
                  f"features max|d| {d_feat:.2e} (|f|max {feat.abs().max():.3f}, cos {cos:.8f})")
            assert d_pre < 1e-5 and d_anon < 1e-4 and d_feat < 2e-4 * max(1.0, feat.abs().max().item())
            feats[tag] = feat
            if tag == "test":
                out[f"{name}/features"] = feat.numpy().astype(np.float32)
                out[f"{name}/clip_sha"] = np.array(sha(clip))
                a = anon.numpy()
                out[f"{name}/anon_stats"] = np.array([a.mean(), a.std(), a.min(), a.max()], dtype=np.float64)
                # 256 fixed sample points of the anonymized frames [16,3,h,w]
                rs = np.random.RandomState(7)
                idx = rs.randint(0, a.size, 256)
                out[f"{name}/anon_samples"] = a.reshape(-1)[idx].astype(np.float32)
                if pred is not None:
                    out[f"{name}/pred"] = pred.numpy().astype(np.float32)
        ctrl = torch.nn.functional.cosine_similarity(feats["test"], feats["control"], dim=0).item()
        out[f"{name}/control_cos"] = np.array(ctrl)
        out[f"{name}/control_features"] = feats["control"].numpy().astype(np.float32)
        print(f"{name}: control cosine between different clips = {ctrl:.4f}; feature norm {feats['test'].norm():.2f}")
    # snippet-index golden vectors of the ShanghaiTech reader, from the reference read_video itself
    out.update(shanghai_index_golden())
    # the x[0,t,c] = 10t+c probe of the anonymizer->encoder raw reshape (dali_extraction.py:171-173)
    probe = torch.zeros(1, 16, 3, 1, 1)
    for t in range(16):
        for c in range(3):
            probe[0, t, c] = 10 * t + c
    ori = probe.permute(0, 2, 1, 3, 4).shape
    out["glue/probe"] = probe.view(-1, 3, 1, 1).reshape(*ori).numpy().reshape(3, 16)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote golden_v1.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
