"""A handful of SLAB-feed launches at bench shapes, for `ncu --set full -k regex:conv_slab` (not a pytest file)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ted-spad_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_diag as G  # noqa: E402
from tedspad_b200 import _lib as L  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = L.SLAB_3X3
G.time_slab("64->64 @224", K, N, (1, 224, 224), 64, 64, (1, 3, 3), iters=1)
G.time_slab("128->64 @224", K, N, (1, 224, 224), 128, 64, (1, 3, 3), iters=1)
G.time_slab("64->128 @112", K, N, (1, 112, 112), 64, 128, (1, 3, 3), iters=1)
G.time_slab("stem2d 3->64 @224", L.SLAB_STEM2D, N, (1, 224, 224), 8, 64, (1, 3, 3), iters=1, cin_real=3)
G.time_slab("stem3d i3d", L.SLAB_STEM3D, max(1, N // 16), (16, 224, 224), 4, 64, (7, 7, 7), stride=(2, 2, 2), pad_f=(2, 2, 2),
            pad_b=(3, 3, 3), iters=1, cin_real=3)
torch.cuda.synchronize()
