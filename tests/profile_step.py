"""Per-launch timing table of one bench step (32 clips, UNet + I3D): CUDA events around every conv launch,
and (under ncu) the launch list.  Usage: python tests/profile_step.py [batch_clips] [encoder arch] [anonymizer arch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))
import bench  # noqa: E402
from tedspad_b200 import ops  # noqa: E402
from tedspad_b200.extraction import SnippetExtractor, crop_boxes  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ARCH = sys.argv[2] if len(sys.argv) > 2 else "i3d"       # e.g. largei3d (the encoder the reference scripts configure)
FA_ARCH = sys.argv[3] if len(sys.argv) > 3 else "unet"   # or unet++ (the anonymizer the reference scripts configure)
dev = torch.device("cuda", 0)
fa, ft = bench.build_models(dev, FA_ARCH, ARCH)
ext = SnippetExtractor(fa, ft, reso=bench.RESO, batch_clips=B)
(ch, cw), boxes = crop_boxes(*bench.SRC_HW)
desc = np.zeros((B * 16, 4), dtype=np.int32)
desc[:, 0] = np.arange(B * 16)
desc[:, 1], desc[:, 2] = boxes[0][0], boxes[0][1]
frames = bench.synthetic_frames(1, B * 16, bench.SRC_HW).to(dev)
for _ in range(2):
    ext.features_of_clips(frames, desc, (ch, cw))
torch.cuda.synchronize()
ops.CONV_EVENTS = []
ops.OP_EVENTS = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly this steady-state step
e0.record()
ext.features_of_clips(frames, desc, (ch, cw))
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
tot = e0.elapsed_time(e1)
rows = []
for a, b, (n, d, h, w, c, co, k, s, od, oh, ow, cin) in ops.CONV_EVENTS:
    ms = a.elapsed_time(b)
    gf = 2.0 * n * od * oh * ow * cin * co * k[0] * k[1] * k[2] / 1e9
    rows.append((ms, gf, f"in[{n},{d},{h},{w},{c}] -> {co} k{k} s{s}"))
conv = sum(r[0] for r in rows)
print(f"step {tot:.2f} ms, conv {conv:.2f} ms over {len(rows)} launches, B={B}")
for ms, gf, name in rows:
    print(f"{ms:8.3f} ms {gf / ms:8.1f} TFLOP/s  {gf:9.1f} GF  {name}")
from collections import defaultdict
agg = defaultdict(lambda: [0, 0.0])
for a, b, name in ops.OP_EVENTS:
    agg[name][0] += 1
    agg[name][1] += a.elapsed_time(b)
print("non-convolution launches (CUDA events, in-step):")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms  n={n:3d}  {name}")
for a, b, name in ops.OP_EVENTS:
    if name in ("upsample2x", "maxpool", "upsample2x_nearest", "frames_to_clip"):
        print(f"   {a.elapsed_time(b):7.3f} ms {name}")
