"""I3D-ResNet50 (`load_ft_model(arch='largei3d')` -> wrapper_i3d.i3d), drop-in for the reference's
aux_code/models/large_i3d.py:130-263: same parameter names (`conv1.weight`, `layer2.0.downsample.1.bias`,
`fc.weight`, ...) and `extract_features(x[B,3,T,H,W]) -> [B,2048,1,1,1]`.  Parameter containers only;
tedspad_b200.engine.I3Res50Executor does the arithmetic (BN, residual add and ReLU fused into the
convolution epilogues)."""
import torch
import torch.nn as nn

from aux_code._base import CudaModule
from tedspad_b200.engine import I3RES50_LAYERS, I3Res50Executor


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample, temp_conv, temp_stride, use_nl=False):
        super().__init__()
        if use_nl or temp_stride != 1:
            raise NotImplementedError("NonLocal blocks / temporal stride are never built by the reference "
                                      "(model_loaders.py:262, large_i3d.py:142-145)")
        self.conv1 = nn.Conv3d(inplanes, planes, (1 + temp_conv * 2, 1, 1), (1, 1, 1), (temp_conv, 0, 0), bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, (1, 3, 3), (1, stride, stride), (0, 1, 1), bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        if downsample is not None:
            self.downsample = downsample


class I3Res50(CudaModule):
    executor_cls = I3Res50Executor
    executor_kwargs = {"prefix": ""}

    def __init__(self, block=Bottleneck, layers=(3, 4, 6, 3), num_classes=400, use_nl=False):
        super().__init__()
        if use_nl or tuple(layers) != (3, 4, 6, 3):
            raise NotImplementedError("only I3Res50(layers=[3,4,6,3], use_nl=False) is built by the reference")
        self.conv1 = nn.Conv3d(3, 64, (5, 7, 7), (2, 2, 2), (2, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        inpl = 64
        for li, (planes, nblocks, stride, tcs) in enumerate(I3RES50_LAYERS, 1):
            blocks = []
            for b in range(nblocks):
                ds = None
                if b == 0:
                    ds = nn.Sequential(nn.Conv3d(inpl, planes * 4, 1, (1, stride, stride), bias=False),
                                       nn.BatchNorm3d(planes * 4))
                blocks.append(Bottleneck(inpl, planes, stride if b == 0 else 1, ds, tcs[b], 1))
                inpl = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.fc = nn.Linear(2048, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')

    def extract_features(self, x):
        ex = self._exec(x)
        with torch.cuda.device(x.device):
            enc = self._to_cl(x)
            feat = self._graphed(ex, ("extract_features",) + tuple(x.shape), lambda: ex.run(enc))  # [B,1,2048]
            return feat.reshape(feat.shape[0], 2048, 1, 1, 1).clone()

    def features_from_cl(self, enc_in):
        return self._exec(enc_in.buf).run(enc_in)

    def forward(self, x):
        """(logits [B, num_classes], feat) as large_i3d.py:229-246 in eval mode (dropout is the identity):
        feat = pooled.squeeze() ([B,2048], or [2048] when B == 1, like the reference's `x.squeeze()`)."""
        ex = self._exec(x)
        with torch.cuda.device(x.device):
            enc = self._to_cl(x)
            feat = self._graphed(ex, ("extract_features",) + tuple(x.shape), lambda: ex.run(enc)).reshape(x.shape[0], 2048).clone()
            pc = self.__dict__.get("_tsp_fc")
            if pc is None or pc[0] != ex.serial:
                from tedspad_b200.ops import PackedConv
                pc = (ex.serial, PackedConv(self.fc.weight, self.fc.bias, None, device=x.device))
                self.__dict__["_tsp_fc"] = pc
            from tedspad_b200 import ops
            logits = ops.linear(feat, pc[1], (ex.bufs, "fc")).clone()
        return logits, feat.squeeze()
