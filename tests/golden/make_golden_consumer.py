"""Generates tests/golden/consumer_v1.npz by running the UNMODIFIED consumer of the feature files,
anomaly_detection_mgfn/datasets/dataset.py `Dataset.__getitem__` (+ utils/utils.py `process_feat`), imported from
/root/reference (this container only).  The module sets the default tensor type to CUDA and parses argv at import
time; both are neutralised here (no GPU in this container), `visdom` is stubbed, and the Dataset object is built
without its list-file parser - `__getitem__` itself runs as written, on feature files this script saves.

Run:  python tests/golden/make_golden_consumer.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import consumer as C  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from _cases import CONSUMER_CASES as CASES, consumer_case_features as case_features  # noqa: E402

REF = "/root/reference/anomaly_detection_mgfn"


def _load(name, rel):
    """Import one reference file under the module name its siblings use (`datasets` / `utils` collide with installed
    packages, so the files are loaded by path)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_dataset():
    _load("option", "option.py")
    sys.modules["utils"] = types.ModuleType("utils")
    _load("utils.utils", "utils/utils.py")
    return _load("mgfn_dataset", "datasets/dataset.py").Dataset


def main():
    sys.modules.setdefault("visdom", types.ModuleType("visdom"))
    torch.set_default_tensor_type = lambda *_: None
    sys.argv = ["make_golden_consumer"]
    Dataset = import_reference_dataset()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name in CASES:
            feats = case_features(name)
            path = os.path.join(d, name + "_ours.npy")
            np.save(path, feats)
            for test_mode in (True, False):
                ds = Dataset.__new__(Dataset)
                ds.list, ds.test_mode, ds.tranform, ds.is_normal = [path.replace("_ours", "_mgfn") + "\n"], test_mode, None, True
                got, _ = ds[0]
                mine = C.getitem_test(C.load_features(path)) if test_mode else C.getitem_train(C.load_features(path))
                assert got.dtype == np.float32 and got.shape == mine.shape and np.array_equal(got, mine), (name, test_mode)
                out[f"{name}/{'test' if test_mode else 'train'}"] = got
                print(f"{name} test_mode={test_mode}: reference {got.shape} == oracle bit for bit")
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "consumer_v1.npz"), **out)


if __name__ == "__main__":
    main()
