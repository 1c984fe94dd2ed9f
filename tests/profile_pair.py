"""Three 64-output-channel SLAB launches at bench shapes (single-CTA and CTA-pair kinds), for
`ncu --set full --import-source on -k regex:conv_slab` (not a pytest file)."""
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
G.time_slab("64->64 @224 single", L.SLAB_3X3, N, (1, 224, 224), 64, 64, (1, 3, 3), iters=1)
G.time_slab("64->64 @224 pair", L.SLAB_3X3_PAIR, N, (1, 224, 224), 64, 64, (1, 3, 3), iters=1)
G.time_slab("128->64 @224 pair", L.SLAB_3X3_PAIR, N, (1, 224, 224), 128, 64, (1, 3, 3), iters=1)
torch.cuda.synchronize()
