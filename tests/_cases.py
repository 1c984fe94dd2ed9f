"""Shared test fixtures: the golden cases (seeds, sizes) and deterministic weights/inputs for them.
Everything here is regenerated from seeds on any box; tests/golden/golden_v1.npz (made by
tests/golden/make_golden.py from the real reference) pins that the regeneration is faithful."""
import functools
import os

import numpy as np
import torch

from oracle import models as M
from oracle import preprocess as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")

CASES = {
    # name: (encoder arch, source (H, W), output reso, weight seeds (fa, ft), clip seeds (calib, test, control))
    "unet_i3d_224": ("i3d", (240, 320), (224, 224), (1, 2), (100, 101, 102)),
    "unet_largei3d_224": ("largei3d", (240, 320), (224, 224), (1, 3), (100, 101, 102)),
    "unet_r3d18_112": ("r3d_18", (120, 160), (112, 112), (4, 5), (110, 111, 112)),
    # the configuration both reference scripts request (dali_extraction.py:122-123): UNet++ anonymizer + I3D-ResNet50.
    # smp is not installable here, so this case has NO golden vectors from the reference: its oracle is the
    # restatement of smp 0.3.3's UnetPlusPlus in oracle/models.py (parity unpinned for the decoder half).
    "unetpp_largei3d_224": ("largei3d", (240, 320), (224, 224), (6, 7), (100, 101, 102)),
}
FA_ARCH = {"unetpp_largei3d_224": "unet++"}   # anonymizer per case (default 'unet')


def fa_arch(name):
    return FA_ARCH.get(name, "unet")

# Synthetic-init regime per encoder (oracle.models.calibrated_state_dict): BN beta/gamma range and the mean the
# calibration features are scaled to.  The residual encoders pass the bf16 gate at beta/gamma in (0.5, 1.5); the
# plain 21-deep Inception stack has no skip paths to damp the random network's perturbation gain and needs the
# more linear (1.5, 2.5) regime.  Its feature scale is 0.10: the absolute gate (2e-2) is scale-dependent, and at
# this scale the bf16 noise floor of the 40-conv stack (measured: ours and cuDNN-bf16 autocast both land at
# max|err| = 0.018 +- 0.004 per unit of 0.15 feature mean, i.e. ON the gate, flipping with any change of rounding
# order) sits at ~0.6 of the gate, so the assertion tests the kernels rather than the noise realisation.
# tests/test_gpu_parity.py also checks, for every regime including the chaotic stress one, that this pipeline
# deviates from fp32 no more than stock PyTorch bf16 autocast (cuDNN) does on the very same network.
INIT = {
    "unet": {"beta_over_gamma": (0.5, 1.5)},
    "unet++": {"beta_over_gamma": (0.5, 1.5)},
    "i3d": {"beta_over_gamma": (1.5, 2.5), "feature_mean": 0.10},
    "largei3d": {"beta_over_gamma": (0.5, 1.5), "feature_mean": 0.5},
    "r3d_18": {"beta_over_gamma": (0.5, 1.5), "feature_mean": 0.5},
}
STRESS_INIT = {"beta_over_gamma": (-0.3, 0.3)}   # SURVEY 8c's first suggestion: chaotic for any 16-bit evaluation


def golden():
    return np.load(GOLDEN)


@functools.lru_cache(maxsize=None)
def case_weights(name, stress=False):
    arch, hw, reso, wseeds, cseeds = CASES[name]
    clip = M.structured_clip_u8(cseeds[0], 16, hw[0], hw[1])
    x = torch.from_numpy(P.dali_val_augmentations(clip, reso))
    fa_init = STRESS_INIT if stress else INIT[fa_arch(name)]
    ft_init = dict(INIT[arch], **STRESS_INIT) if stress else INIT[arch]
    with torch.no_grad():
        sd_fa = M.calibrated_state_dict(fa_arch(name), wseeds[0], x, **fa_init)
        enc_in = M.anonymize_and_reshape(sd_fa, x.unsqueeze(0))
        sd_ft = M.calibrated_state_dict(arch, wseeds[1], enc_in, **ft_init)
    return sd_fa, sd_ft


def case_clip(name, which="test"):
    arch, hw, reso, wseeds, cseeds = CASES[name]
    seed = {"calib": cseeds[0], "test": cseeds[1], "control": cseeds[2]}[which]
    return M.structured_clip_u8(seed, 16, hw[0], hw[1])


def oracle_features(name, clip_u8, stress=False):
    """fp32 oracle: uint8 frames -> (preprocessed [16,3,h,w], anonymized enc_in [1,3,16,h,w], features [F])."""
    arch, hw, reso, _, _ = CASES[name]
    sd_fa, sd_ft = case_weights(name, stress)
    x = torch.from_numpy(P.dali_val_augmentations(clip_u8, reso))
    with torch.no_grad():
        enc_in = M.anonymize_and_reshape(sd_fa, x.unsqueeze(0))
        feat = M.encoder_features(arch, sd_ft, enc_in).squeeze(0)
    return x, enc_in, feat


def parity_metrics(got, ref):
    got, ref = got.double().flatten(), ref.double().flatten()
    cos = float(torch.nn.functional.cosine_similarity(got, ref, dim=0))
    gc, rc = got - got.mean(), ref - ref.mean()
    ccos = float(torch.nn.functional.cosine_similarity(gc, rc, dim=0))
    return {"cos": cos, "centered_cos": ccos, "max_abs": float((got - ref).abs().max()),
            "ref_max": float(ref.abs().max())}


# ---- consumer contract (SURVEY a-12): feature matrices as the extraction scripts write them (float64), loaded the way
# anomaly_detection_mgfn/datasets/dataset.py does.  tests/golden/consumer_v1.npz holds the reference's own outputs.
CONSUMER_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consumer_v1.npz")
CONSUMER_CASES = {"t7_f8": (7, None, 8), "t40_f8": (40, None, 8), "t75_c10_f6": (75, 10, 6), "t32_c5_f4": (32, 5, 4),
                  "t33_f3": (33, None, 3)}   # name: (T, ncrops | None, F)


def consumer_case_features(name):
    T, ncrops, F = CONSUMER_CASES[name]
    rs = np.random.RandomState(sum(map(ord, name)))
    shape = (T, F) if ncrops is None else (T, ncrops, F)
    return rs.rand(*shape).astype(np.float64) * 2.0   # the extraction scripts write float64 (dali_extraction.py:163)
