"""Device-side consumer view of extracted features: what anomaly_detection_mgfn's
`Dataset.__getitem__` (datasets/dataset.py:51-132) builds on the host from a loaded `.npy` - float32 cast,
`[T,F] -> [T,1,F]`, per-crop `process_feat` resampling to 32 segments (utils/utils.py:34-42), L2 magnitude appended
as feature F+1 - computed straight from the feature rows still resident on the GPU, so that online anomaly scoring
does not need the file round trip.  The arithmetic is tedspad_mgfn_rows (csrc/ops.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def segment_bounds(T, seg_length=32):
    """utils/utils.py:36: np.linspace(0, len(feat), length + 1, dtype=int)."""
    return np.linspace(0, T, seg_length + 1, dtype=int).astype(np.int32)


def _as_t_c_f(features):
    if not (isinstance(features, torch.Tensor) and features.is_cuda):
        raise RuntimeError("mgfn consumer: a CUDA tensor is required - this framework has no CPU path "
                           "(on the host, anomaly_detection_mgfn/datasets/dataset.py does this itself)")
    f = features.to(torch.float32)                 # dataset.py:55 np.array(features, dtype=np.float32)
    if f.dim() < 3:
        f = f.unsqueeze(1)                         # dataset.py:70-71 / 87-88 expand_dims(axis=1)
    return f.contiguous()


def getitem_test(features):
    """dataset.py:68-86 (test_mode): [T,F] or [T,ncrops,F] -> float32 cuda [T, ncrops, F+1]."""
    f = _as_t_c_f(features)
    T, nc, F_ = f.shape
    out = torch.empty((T, nc, F_ + 1), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        L.check(L.lib().tedspad_mgfn_rows(f.data_ptr(), T, nc, F_, None, 0, 0, out.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tedspad_mgfn_rows")
    return out


def getitem_train(features, seg_length=32):
    """dataset.py:87-99 (training): -> float32 cuda [ncrops, seg_length, F+1]."""
    f = _as_t_c_f(features)
    T, nc, F_ = f.shape
    bounds = torch.from_numpy(segment_bounds(T, seg_length)).to(f.device)
    out = torch.empty((nc, seg_length, F_ + 1), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        L.check(L.lib().tedspad_mgfn_rows(f.data_ptr(), T, nc, F_, bounds.data_ptr(), seg_length, 1, out.data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tedspad_mgfn_rows")
    return out
