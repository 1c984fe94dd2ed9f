"""Bit-level fingerprint of tedspad_upsample2x on fixed inputs (compare across TEDSPAD_UP_V settings: the forms of the
kernel must agree bit for bit) + its time on the four UNet levels of a 32-clip step."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ted-spad_b200")]
import torch
from tedspad_b200 import ops

torch.manual_seed(0)
h = hashlib.sha256()
for (n, c, hh, ww), (th, tw), ld, coff in [((3, 64, 7, 7), (14, 14), 128, 64), ((2, 128, 6, 10), (13, 21), 256, 128),
                                            ((4, 64, 56, 56), (112, 112), 128, 64), ((2, 512, 14, 14), (28, 28), 1024, 512),
                                            ((2, 64, 33, 17), (67, 35), 64, 0)]:
    x = torch.randn(n, c, hh, ww, device="cuda")
    xc = ops.CLTensor(n, 1, hh, ww, c, (0, 1, 1), device="cuda")
    xc.buf.zero_()
    xc.interior()[...] = x.permute(0, 2, 3, 1).unsqueeze(1).to(torch.bfloat16)
    cat = ops.CLTensor(n, 1, th, tw, ld, (0, 1, 1), device="cuda")
    cat.buf.fill_(3.0)
    ops.upsample2x(xc, cat.slice(coff, c))
    torch.cuda.synchronize()
    h.update(cat.buf.view(torch.int16).cpu().numpy().tobytes())
print("TEDSPAD_UP_V=%s sha256 %s" % (os.environ.get("TEDSPAD_UP_V", "default"), h.hexdigest()[:24]))

tot = 0.0
for c, s in ((512, 14), (256, 28), (128, 56), (64, 112)):
    n = 512
    xc = ops.CLTensor(n, 1, s, s, c, (0, 1, 1), device="cuda")
    xc.buf.normal_()
    cat = ops.CLTensor(n, 1, 2 * s, 2 * s, 2 * c, (0, 1, 1), device="cuda")
    y = cat.slice(c, c)
    for _ in range(3):
        ops.upsample2x(xc, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.upsample2x(xc, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    wr = n * 4 * s * s * c * 2 / 1e9
    print("  %4d ch %3d^2 -> %3d^2 x %d frames: %.3f ms  (%.2f GB written, %.2f TB/s read + write)" % (c, s, 2 * s, n, ms, wr, wr * 1.25 / ms))
    del xc, cat
print("  sum %.3f ms" % tot)
