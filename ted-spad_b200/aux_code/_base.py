"""Shared behaviour of the boundary nn.Modules: parameters live in torch (reference state_dict
names), arithmetic lives in libtedspad.so; the packed-weight executor is rebuilt whenever the
parameters change (load_state_dict, .cuda(), .to())."""
import os
import sys

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tedspad_b200 import ops  # noqa: E402
from tedspad_b200.ops import CLTensor  # noqa: E402


class CudaModule(nn.Module):
    """nn.Module whose forward runs on the B200 kernels only (no CPU / eager fallback)."""

    executor_cls = None
    executor_kwargs = {}
    _serial = 0

    def _replicate_for_data_parallel(self):
        # nn.DataParallel (dali_extraction.py:126-141) broadcasts the parameters to per-device replicas whose
        # parameters() are empty and calls forward from one thread per device.  The reference's own multi-GPU branch
        # cannot run (its DataParallel-wrapped ft_model has no .extract_features / .i3d); this
        # implementation scales as one process per GPU over a sharded video list instead.
        raise RuntimeError(f"{type(self).__name__}: nn.DataParallel is not supported - run one process per GPU "
                           "(torchrun) and shard the video list with tedspad_b200.extraction.extract_dataset_distributed")

    def __setattr__(self, name, value):
        if isinstance(value, (nn.Module, nn.Parameter)):   # e.g. model.i3d.fc = nn.Linear(...) (model_loaders.py:194-195)
            self.__dict__.pop("_tsp_tensors", None)
        super().__setattr__(name, value)

    def _apply(self, fn, *a, **k):
        # .cuda() / .to() / .float(): parameters may be re-created - drop the cached tensor list
        self.__dict__.pop("_tsp_tensors", None)
        return super()._apply(fn, *a, **k)

    def _signature(self):
        """Identity + in-place version of every parameter and buffer.  The tensor list is cached (walking the module
        tree through state_dict() cost 0.5-0.7 ms per call, twice per clip in the batch-1 drop-in loop); it is rebuilt
        whenever the number of registered tensors changes or _apply ran, and `_version` catches load_state_dict /
        optimizer-style in-place updates."""
        d = self.__dict__
        ts = d.get("_tsp_tensors")
        d["_tsp_calls"] = d.get("_tsp_calls", 0) + 1
        if ts is None or d["_tsp_calls"] % 256 == 0:   # (periodic full walk: catches a swapped grand-child module)
            ts = [v for _, v in self.state_dict(keep_vars=True).items()]
            d["_tsp_tensors"] = ts
        return tuple((v.data_ptr(), v._version) for v in ts) + (str(ts[0].device) if ts else "",)

    def _exec(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError(f"{type(self).__name__}: input must be a CUDA tensor - this implementation runs only "
                               "on the sm_100a kernels in libtedspad.so and has no CPU fallback")
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: inference only (BatchNorm uses running statistics); "
                               "call .eval() first as dali_extraction.py:139-141 does")
        p = next(self.parameters())
        if p.device != x.device:
            raise RuntimeError(f"{type(self).__name__}: parameters on {p.device}, input on {x.device}")
        sig = self._signature()
        if getattr(self, "_tsp_sig", None) != sig:
            with torch.cuda.device(x.device):
                ex = self.executor_cls(self.state_dict(), x.device, **self.executor_kwargs)
            CudaModule._serial += 1
            ex.serial = CudaModule._serial      # identity of this weight snapshot (keys captured CUDA graphs)
            self.__dict__["_tsp_executor"] = ex
            self.__dict__["_tsp_sig"] = sig
        return self.__dict__["_tsp_executor"]

    def executor(self, device):
        """The packed-weight executor for the current parameters on `device` (rebuilt when they changed)."""
        return self._exec(torch.empty(1, device=device))

    def _graphed(self, ex, key, fn):
        """fn() over the executor's static buffers: eager on first use, a replayed CUDA graph afterwards."""
        return ex.graphs.run(key, ex.bufs.generation, fn)

    def _to_cl(self, x, name="in"):
        """fp32 [B,3,T,H,W] / [N,3,H,W] -> channels-last bf16 [.., 4|8] (pad channels zero)."""
        ex = self.__dict__["_tsp_executor"]
        if x.dim() == 4:
            n, c, h, w = x.shape
            t = 1
        else:
            n, c, t, h, w = x.shape
        if c != 3:
            raise RuntimeError(f"{type(self).__name__}: expected 3 input channels, got {c}")
        from tedspad_b200 import engine
        # clips feed an encoder stem (4-channel pixels for the SLAB stem); frames feed the UNet (8-channel pixels)
        buf = ex.bufs.get(name, n, t, h, w, engine.ENC_IN_CHANNELS if x.dim() == 5 else 8)
        return ops.nchw_to_cl(x, buf)
