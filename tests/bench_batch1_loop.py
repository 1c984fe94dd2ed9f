"""The reference's batch-1 loop (dali_extraction.py:168-179 verbatim, bench.dropin_batch1_clips_per_s) on the three
anonymizer + encoder pairs, 3 x 96 clips each - a steadier reading than the 64-clip figure inside bench.py
(not a pytest file).  Usage: python tests/bench_batch1_loop.py"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'ted-spad_b200'))
import bench
dev = torch.device('cuda', 0)
for fa_arch, enc in (("unet++", "largei3d"), ("unet", "largei3d"), ("unet", "i3d")):
    fa, ft = bench.build_models(dev, fa_arch, enc)
    r = [bench.dropin_batch1_clips_per_s(fa, ft, dev, n_clips=96)["value"] for _ in range(3)]
    print(os.environ.get("TEDSPAD_STEM_PAIR_MIN_KD"), os.environ.get("TEDSPAD_STEM_PAIR"), fa_arch, enc, [round(v, 1) for v in r], flush=True)
