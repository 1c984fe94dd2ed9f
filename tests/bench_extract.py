"""Dataset-shaped end-to-end run of the extraction driver (not a pytest file): synthetic decoded videos in pinned host
memory -> SnippetExtractor.extract_video (copy-stream prefetch, N-crop) -> float64 .npy files, the way
dali_extraction.py / st_feature_extraction.py are used.  Shapes follow BASELINE.json configs[2] (UCF-Crime: 320x240,
10-crop) and configs[3] (ShanghaiTech: 480x856 -> 224, single crop, PIL resampling), scaled down in video count.

    python tests/bench_extract.py [n_videos] [frames_per_video] [arch]
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))
import bench  # noqa: E402
from tedspad_b200.extraction import SnippetExtractor, extract_dataset  # noqa: E402

n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
arch = sys.argv[3] if len(sys.argv) > 3 else "i3d"
dev = torch.device("cuda", 0)
fa, ft = bench.build_models(dev)
if arch != "i3d":
    from aux_code.model_loaders import load_ft_model
    ft = load_ft_model(arch=arch, num_classes=102).to(dev).eval()


def videos(h, w, n, frames, step=37):
    """n videos of ~`frames` frames each, all windows of ONE pinned pool of random frames (so that a run of several
    seconds does not need hundreds of GB of host memory; the frame bytes still cross PCIe for every video)."""
    g = torch.Generator().manual_seed(7)
    longest = frames + step * (n - 1)
    pool = torch.randint(0, 256, (longest, h, w, 3), generator=g, dtype=torch.uint8).pin_memory()
    return [(f"/data/video_{i:03d}.mp4", frames + step * i, (lambda i=i: pool[:frames + step * i])) for i in range(n)]


def run(name, ext, vids):
    with tempfile.TemporaryDirectory() as d:
        extract_dataset(ext, vids[:1], os.path.join(d, "warm"), log=lambda *_: None)      # warm-up (buffers, packing)
        torch.cuda.synchronize()
        h2d0 = ext.h2d_bytes
        t0 = time.perf_counter()
        written = extract_dataset(ext, vids, os.path.join(d, "out"), log=lambda *_: None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = ext.h2d_bytes - h2d0
        rows = [np.load(f) for f in written]
    snips = sum(r.shape[0] for r in rows)
    clips = snips * ext.ncrops
    frames = sum(v[1] for v in vids)
    print(f"{name}: {len(vids)} videos, {frames} frames, {snips} snippets x {ext.ncrops} crops = {clips} clip forwards in "
          f"{dt:.2f} s -> {clips / dt:.1f} clips/s, {frames / dt:.0f} source frames/s; host->device {h2d / 1e9:.2f} GB = "
          f"{h2d / dt / 1e9:.1f} GB/s ({h2d / max(clips, 1) / 1e6:.1f} MB per clip forward); row shape {rows[0].shape} {rows[0].dtype}",
          flush=True)


run("UCF-Crime-shaped 10-crop (configs[2])", SnippetExtractor(fa, ft, source="dali", ncrops=10, batch_clips=40),
    videos(240, 320, n_videos, n_frames))
run("UCF-Crime-shaped single crop", SnippetExtractor(fa, ft, source="dali", ncrops=1, batch_clips=32),
    videos(240, 320, n_videos, n_frames))
run("ShanghaiTech-shaped single crop, PIL path (configs[3])", SnippetExtractor(fa, ft, source="shanghai", ncrops=1, batch_clips=32),
    videos(480, 856, int(os.environ.get("BENCH_ST_VIDEOS", n_videos * 4)), max(64, n_frames // 4), step=3))
