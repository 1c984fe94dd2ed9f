"""Anonymizer-only streaming: what visualization/visualize_anonymization.py:65-115 does with the
reference modules - every frame of a video at its NATIVE resolution through `fa_model`, colour-flipped, min-max
normalised to uint8 - on the B200 kernels.  Frames go from uint8 HWC to the anonymizer's channels-last bf16 input with
the preprocessing kernel (crop = the whole frame, identity resampling = `ToPILImage -> ToTensor`, :70-73,95-96), in
chunks, so that a long video does not need all its activations at once (the reference stacks the whole video into one
batch, :98-104)."""
import numpy as np
import torch

from . import _lib as L
from . import ops


def anonymize_frames(fa_model, frames_u8, chunk=16, flip_channels=True):
    """frames_u8: uint8 [F,H,W,3] (CPU or CUDA) -> float32 CUDA [F,3,H,W]: `fa_model(frames / 255)` per frame, with
    `torch.flip(output, dims=[1])` applied like visualize_anonymization.py:103-104 when flip_channels."""
    dev = next(fa_model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("anonymize_frames needs a CUDA anonymizer: there is no CPU path")
    F_, H, W, _ = frames_u8.shape
    out = torch.empty((F_, 3, H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ex = fa_model.executor(dev)
        for f0 in range(0, F_, chunk):
            fr = frames_u8[f0:f0 + chunk].to(dev, non_blocking=True).contiguous()
            n = fr.shape[0]
            desc = torch.zeros((n, 4), dtype=torch.int32)
            desc[:, 0] = torch.arange(n, dtype=torch.int32)
            x0 = ex.input_buffer(n, H, W)
            ops.preprocess(fr, desc.to(dev), (H, W), x0, L.RESAMPLE_AA_FLOAT)      # scale 1: identity taps, /255, bf16
            clip = ex.bufs.get("vis_clip", n, 1, H, W, 4)
            ex.run(x0, clip, T=1, frames_out=out[f0:f0 + n])
    return torch.flip(out, dims=[1]) if flip_channels else out


def to_uint8_video(tensor):
    """save_video's normalisation (visualize_anonymization.py:50-58): [T,3,H,W] -> uint8 numpy [T,H,W,3], min-max over
    the WHOLE tensor."""
    t = tensor.permute(0, 2, 3, 1).cpu().numpy()
    t = (t - t.min()) / (t.max() - t.min())
    return (t * 255).astype(np.uint8)


def save_video_cv2(frames_u8_thwc, filename, fps=30.0):
    """imageio (the reference's writer) is not in this image; cv2's MJPG writer takes the same uint8 [T,H,W,3] frames."""
    import cv2
    T, H, W, _ = frames_u8_thwc.shape
    wr = cv2.VideoWriter(filename, cv2.VideoWriter_fourcc(*"MJPG"), float(fps), (W, H))
    for fr in frames_u8_thwc:
        wr.write(np.ascontiguousarray(fr))
    wr.release()
