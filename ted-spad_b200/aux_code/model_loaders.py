"""Drop-in for the reference's aux_code/model_loaders.py on the feature-extraction path:
`load_fa_model` (:17-52) and `load_ft_model` (:56-90) keep their signatures, checkpoint formats
(`fa_model_state_dict` / `ft_model_state_dict`, the 'module.' prefix strip, the FrozenBN 'scale'->'weight'
rename and the `.i3d` fallback) and return nn.Modules whose parameter names equal the reference's,
so reference checkpoints load with strict=True.  The returned modules compute with the sm_100a
kernels of libtedspad.so; there is no CPU path."""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from aux_code._base import CudaModule
from aux_code.models.i3d import InceptionI3d
from aux_code.models.large_i3d import I3Res50
from aux_code.models.unet_model import UNet
from aux_code.models.unetpp import UnetPlusPlus
from tedspad_b200.engine import R3D18Executor


def _load(path):
    # reference checkpoints pickle optimizer / GradScaler state next to the weights (train_action.py:391-396)
    return torch.load(path, map_location="cpu", weights_only=False)


def load_fa_model(saved_model_file=None, arch='unet++'):
    if arch == 'unet':
        fa_model = UNet(n_channels=3, n_classes=3)
    elif arch == 'unet++':
        # the same call the reference makes (model_loaders.py:19-30) on the drop-in for smp 0.3.3's UnetPlusPlus
        fa_model = UnetPlusPlus(
            encoder_name='resnet18',
            encoder_depth=4,
            encoder_weights="imagenet" if not saved_model_file else None,   # never downloaded: see UnetPlusPlus
            decoder_channels=(256, 128, 64, 32),
            decoder_attention_type=None,
            decoder_use_batchnorm=True,
            in_channels=3,
            classes=3,
            activation=None,
            aux_params=None
        )
    else:
        raise ValueError(f"Architecture {arch} invalid for fa_model. Try 'unet' or 'unet++'")
    if saved_model_file:
        saved_dict = _load(saved_model_file)
        try:
            fa_model.load_state_dict(saved_dict['fa_model_state_dict'], strict=True)
        except RuntimeError:
            stripped = OrderedDict((k[7:], v) for k, v in saved_dict['fa_model_state_dict'].items())  # 'module.'
            fa_model.load_state_dict(stripped, strict=True)
        print(f'fa_model loaded from {saved_model_file} successfully!')
    else:
        print('fa_model freshly initialized!')
    return fa_model


def load_ft_model(arch='r3d', saved_model_file=None, num_classes=400, kin_pretrained=False):
    if arch == 'i3d':
        ft_model = build_i3d_classifier(num_classes=num_classes, pretrained=kin_pretrained)
    elif arch == 'largei3d':
        ft_model = build_largei3d_classifier(num_classes=num_classes, pretrained=kin_pretrained)
    elif arch == 'r3d_18':
        ft_model = wrapper_r3d_18(num_classes=num_classes, pretrained=kin_pretrained)
    elif arch == 'mvitv2':
        raise NotImplementedError("arch='mvitv2' (model_loaders.py:217-231) is not used by feature extraction")
    else:
        print(f"Architecture {arch} invalid for ft_model. Try 'i3d', 'largei3d', 'mvitv2', or 'r3d_18'.")
        return
    if saved_model_file:
        saved_dict = _load(saved_model_file)
        try:
            ft_model.load_state_dict(saved_dict['ft_model_state_dict'], strict=True)
        except RuntimeError:
            try:
                renamed = OrderedDict((k.replace('scale', 'weight'), v)  # FrozenBN checkpoints (large_i3d.py:16-20)
                                      for k, v in saved_dict['ft_model_state_dict'].items())
                ft_model.load_state_dict(renamed, strict=True)
            except RuntimeError:
                ft_model.i3d.load_state_dict(saved_dict['ft_model_state_dict'], strict=True)
        print(f'ft_model loaded from {saved_model_file} successfully!')
    else:
        print(f'ft_model freshly initialized! Pretrained: {kin_pretrained}')
    return ft_model


def build_i3d_classifier(num_classes=400, pretrained=True):
    temp_classes = 0
    if pretrained:
        temp_classes, num_classes = num_classes, 400
    model = InceptionI3d(num_classes=num_classes, dropout_keep_prob=0.5)
    if pretrained:
        model.load_state_dict(_load(os.path.join('..', 'saved_models', 'rgb_imagenet.pt')), strict=True)
        if temp_classes != 400:
            model.replace_logits(temp_classes)
    return model


def build_largei3d_classifier(num_classes=400, pretrained=True):
    temp_classes = 0
    if pretrained:
        temp_classes, num_classes = num_classes, 400
    model = wrapper_i3d(num_classes=num_classes)
    if pretrained:
        model.i3d.load_state_dict(_load(os.path.join('..', 'saved_models', 'i3d_r50_kinetics.pth')), strict=True)
        if temp_classes != 400:
            model.i3d.fc = nn.Linear(512 * 4, temp_classes)
    return model


class mlp(nn.Module):
    """The embedding head (model_loaders.py:235-254): fc1 + bn1 + ReLU, fc2 + bn2, L2 normalisation.  Parameters
    under the reference's names so that wrapper_i3d checkpoints load strict=True; not on the extraction path (the
    scripts call .i3d.extract_features), but wrapper_i3d.forward uses it."""

    def __init__(self, final_embedding_size=128, use_normalization=True):
        super().__init__()
        self.final_embedding_size, self.use_normalization = final_embedding_size, use_normalization
        self.fc1 = nn.Linear(2048, 512, bias=True)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(128)
        self.fc2 = nn.Linear(512, final_embedding_size, bias=False)

    def forward(self, x, bufs):
        """x: fp32 cuda [B, 2048] -> [B, 128], eval mode (BatchNorm1d folded into the two linear layers)."""
        from tedspad_b200 import _lib as L, ops
        from tedspad_b200.ops import PackedConv
        sig = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if self.__dict__.get("_tsp_sig") != sig:
            bn = lambda m: (m.weight, m.bias, m.running_mean, m.running_var, m.eps)  # noqa: E731
            self.__dict__["_tsp_pcs"] = (PackedConv(self.fc1.weight, self.fc1.bias, bn(self.bn1), device=x.device),
                                         PackedConv(self.fc2.weight, None, bn(self.bn2), device=x.device))
            self.__dict__["_tsp_sig"] = sig
        p1, p2 = self.__dict__["_tsp_pcs"]
        h = ops.linear(x, p1, (bufs, "mlp.fc1"), act=L.ACT_RELU)
        e = ops.linear(h, p2, (bufs, "mlp.fc2")).clone()
        return ops.l2_normalize_rows(e)


class wrapper_i3d(nn.Module):
    """model_loaders.py:258-268.  Deliberately has no `extract_features`, exactly like the reference: the
    extraction scripts reach the encoder through `ft_model.i3d.extract_features` (dali_extraction.py:175-178)."""

    def __init__(self, num_classes=102):
        super().__init__()
        self.i3d = I3Res50(num_classes=num_classes, use_nl=False)
        self.mlp = mlp()

    def forward(self, x):
        """(pred [B, num_classes], feature [B, 128] L2-normalised) as model_loaders.py:265-268, eval mode."""
        if self.training:
            raise RuntimeError("wrapper_i3d: inference only; call .eval() first")
        pred, feature = self.i3d(x)
        with torch.cuda.device(x.device):
            feature = self.mlp(feature.reshape(x.shape[0], 2048), self.i3d.executor(x.device).bufs)
        return pred, feature


def _r3d_trunk():
    """torchvision r3d_18 parameter tree (video/resnet.py:173-181,87-121,198-290) without its arithmetic."""
    conv_bn = lambda ci, co, k, s, p: [nn.Conv3d(ci, co, k, s, p, bias=False), nn.BatchNorm3d(co)]  # noqa: E731

    class Block(nn.Module):
        def __init__(self, ci, co, stride):
            super().__init__()
            self.conv1 = nn.Sequential(*conv_bn(ci, co, 3, stride, 1), nn.ReLU(inplace=True))
            self.conv2 = nn.Sequential(*conv_bn(co, co, 3, 1, 1))
            if stride != 1 or ci != co:
                self.downsample = nn.Sequential(*conv_bn(ci, co, 1, stride, 0))

    trunk = nn.Module()
    trunk.stem = nn.Sequential(*conv_bn(3, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3)), nn.ReLU(inplace=True))
    ci = 64
    for li, co in enumerate([64, 128, 256, 512], 1):
        setattr(trunk, f"layer{li}", nn.Sequential(Block(ci, co, 1 if li == 1 else 2), Block(co, co, 1)))
        ci = co
    trunk.fc = nn.Identity()
    return trunk


class wrapper_r3d_18(CudaModule):
    """model_loaders.py:200-213: forward(x[B,3,T,H,W]) -> (pred[B,num_classes], feature[B,512])."""
    executor_cls = R3D18Executor

    def __init__(self, num_classes=400, pretrained=True):
        super().__init__()
        if pretrained:
            raise NotImplementedError("Kinetics weights for r3d_18 are a torchvision download (no network here); "
                                      "load a checkpoint through load_ft_model(saved_model_file=...) instead")
        self.backbone = _r3d_trunk()
        self.fc = nn.Linear(512, num_classes)

    def forward(self, x):
        ex = self._exec(x)
        with torch.cuda.device(x.device):
            enc = self._to_cl(x)
            pred, feat = self._graphed(ex, ("forward",) + tuple(x.shape), lambda: ex.run(enc))
            return pred.clone(), feat.clone()

    def features_from_cl(self, enc_in):
        return self._exec(enc_in.buf).run(enc_in)[1].unsqueeze(1)
