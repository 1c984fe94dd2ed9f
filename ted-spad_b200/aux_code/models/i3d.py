"""Inception-v1 I3D (`load_ft_model(arch='i3d')`), drop-in for the reference's
aux_code/models/i3d.py:152-340: same constructor arguments, same parameter names
(`logits.conv3d.*` first, then `Conv3d_1a_7x7.conv3d.weight`, `Mixed_3b.b1b.bn.running_mean`, ...)
and the same `extract_features(x[B,3,T,224,224]) -> [B,1024,T/8-1,1,1]`.  The modules below only
hold parameters; tedspad_b200.engine.I3DExecutor does the arithmetic."""
import torch
import torch.nn as nn

from aux_code._base import CudaModule
from tedspad_b200 import ops
from tedspad_b200.engine import I3D_MIXED, I3DExecutor
from tedspad_b200.ops import PackedConv


class Unit3D(nn.Module):
    def __init__(self, in_channels, output_channels, kernel_shape=(1, 1, 1), stride=(1, 1, 1), use_batch_norm=True,
                 use_bias=False):
        super().__init__()
        self.conv3d = nn.Conv3d(in_channels, output_channels, tuple(kernel_shape), tuple(stride), padding=0,
                                bias=use_bias)
        if use_batch_norm:
            self.bn = nn.BatchNorm3d(output_channels, eps=0.001, momentum=0.01)


class InceptionModule(nn.Module):
    def __init__(self, in_channels, oc):
        super().__init__()
        self.b0 = Unit3D(in_channels, oc[0])
        self.b1a = Unit3D(in_channels, oc[1])
        self.b1b = Unit3D(oc[1], oc[2], (3, 3, 3))
        self.b2a = Unit3D(in_channels, oc[3])
        self.b2b = Unit3D(oc[3], oc[4], (3, 3, 3))
        self.b3b = Unit3D(in_channels, oc[5])


class InceptionI3d(CudaModule):
    executor_cls = I3DExecutor
    VALID_ENDPOINTS = ('Conv3d_1a_7x7', 'MaxPool3d_2a_3x3', 'Conv3d_2b_1x1', 'Conv3d_2c_3x3', 'MaxPool3d_3a_3x3',
                       'Mixed_3b', 'Mixed_3c', 'MaxPool3d_4a_3x3', 'Mixed_4b', 'Mixed_4c', 'Mixed_4d', 'Mixed_4e',
                       'Mixed_4f', 'MaxPool3d_5a_2x2', 'Mixed_5b', 'Mixed_5c', 'Logits', 'Predictions')

    def __init__(self, num_classes=400, spatial_squeeze=True, final_endpoint='Logits', name='inception_i3d',
                 in_channels=3, dropout_keep_prob=0.5):
        super().__init__()
        if final_endpoint != 'Logits' or in_channels != 3:
            raise NotImplementedError("only the full RGB network (final_endpoint='Logits') is on the extraction path")
        self._num_classes, self._spatial_squeeze = num_classes, spatial_squeeze
        self.logits = Unit3D(1024, num_classes, use_batch_norm=False, use_bias=True)  # registered first (i3d.py:298)
        self.Conv3d_1a_7x7 = Unit3D(3, 64, (7, 7, 7), (2, 2, 2))
        self.Conv3d_2b_1x1 = Unit3D(64, 64)
        self.Conv3d_2c_3x3 = Unit3D(64, 192, (3, 3, 3))
        for n, cin, oc in I3D_MIXED:
            setattr(self, n, InceptionModule(cin, oc))

    def replace_logits(self, num_classes):
        self._num_classes = num_classes
        self.logits = Unit3D(1024, num_classes, use_batch_norm=False, use_bias=True)

    def extract_features(self, x):
        ex = self._exec(x)
        with torch.cuda.device(x.device):
            enc = self._to_cl(x)
            feat = self._graphed(ex, ("extract_features",) + tuple(x.shape), lambda: ex.run(enc))   # [B, T', 1024] fp32
            return feat.permute(0, 2, 1).reshape(feat.shape[0], 1024, feat.shape[1], 1, 1).clone()

    def features_from_cl(self, enc_in):
        return self._exec(enc_in.buf).run(enc_in)

    def forward(self, x):
        """logits [B, num_classes] (i3d.py:324-333): global average pool + 1x1x1 conv with bias."""
        ex = self._exec(x)
        with torch.cuda.device(x.device):
            enc = self._to_cl(x)
            feat_map = ex.run_trunk(enc)                # Mixed_5c, then AdaptiveAvgPool3d(1)
            pooled = ops.avgpool_features(feat_map, 0)  # [B,1,1024]
            fin = ex.bufs.get("logits_in", x.shape[0], 1, 1, 1, 1024)
            ops.nchw_to_cl(pooled.reshape(x.shape[0], 1024, 1, 1, 1), fin)     # fp32 -> bf16 operand (own kernel)
            pc = PackedConv(self.logits.conv3d.weight, self.logits.conv3d.bias, None, device=x.device)
            out = ex.bufs.get("logits_out", x.shape[0], 1, 1, 1, self._num_classes, dtype=torch.float32)
            ops.conv_forward(fin, pc, out, act=0, y_fp32=True)
        return out.buf.reshape(x.shape[0], self._num_classes).clone()
