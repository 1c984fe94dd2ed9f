"""UNet++ anonymizer (`load_fa_model(arch='unet++')`, the reference's default: aux_code/model_loaders.py:17-30),
drop-in for `segmentation_models_pytorch.UnetPlusPlus(encoder_name='resnet18', encoder_depth=4,
decoder_channels=(256, 128, 64, 32), decoder_use_batchnorm=True, decoder_attention_type=None, in_channels=3,
classes=3, activation=None)` of smp 0.3.3 (third-party, pinned in the reference's pip_requirements.txt:65; its source
is not part of the reference tree).

The modules below only hold parameters, under exactly the names smp gives them, so that the released
`fa_model_state_dict` checkpoints load `strict=True` (incl. the 'module.' strip of model_loaders.py:43-46):

  encoder.conv1.weight, encoder.bn1.*, encoder.layer{1..4}.{0,1}.{conv1,bn1,conv2,bn2}.*, encoder.layer{2..4}.0.downsample.{0,1}.*
      (smp's ResNetEncoder is torchvision's ResNet minus fc/avgpool: layer4 stays in the state_dict although
       encoder_depth=4 never runs it)
  decoder.blocks.x_{d}_{l}.conv{1,2}.0.weight (Conv2d 3x3, bias=False), .conv{1,2}.1.* (BatchNorm2d)
  segmentation_head.0.weight / .bias (Conv2d 3x3 32 -> 3)

`forward(x[N,3,H,W]) -> [N,3,H,W]` (unbounded, activation=None), H and W multiples of 16 as smp's check_input_shape
demands.  tedspad_b200.engine.UNetPPExecutor does the arithmetic on the tcgen05 convolution kernels."""
import warnings

import torch
import torch.nn as nn

from aux_code._base import CudaModule
from tedspad_b200.engine import RESNET18_LAYERS, UNETPP_BLOCKS, UNetPPExecutor


class _BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))


class _ResNet18Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inpl = 64
        for li, (planes, stride) in enumerate(RESNET18_LAYERS, 1):
            setattr(self, f"layer{li}", nn.Sequential(_BasicBlock(inpl, planes, stride), _BasicBlock(planes, planes, 1)))
            inpl = planes


def _conv2d_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _conv2d_relu(cin + cskip, cout)
        self.conv2 = _conv2d_relu(cout, cout)


class _Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.blocks = nn.ModuleDict({name: _DecoderBlock(cin, cskip, cout) for name, cin, cskip, cout in UNETPP_BLOCKS})


class UnetPlusPlus(CudaModule):
    executor_cls = UNetPPExecutor

    def __init__(self, encoder_name='resnet18', encoder_depth=4, encoder_weights='imagenet', decoder_use_batchnorm=True,
                 decoder_channels=(256, 128, 64, 32), decoder_attention_type=None, in_channels=3, classes=3,
                 activation=None, aux_params=None):
        super().__init__()
        if (encoder_name, encoder_depth, tuple(decoder_channels), decoder_use_batchnorm, decoder_attention_type,
                in_channels, classes, activation, aux_params) != ('resnet18', 4, (256, 128, 64, 32), True, None, 3, 3, None, None):
            raise NotImplementedError("only the configuration model_loaders.py:19-30 builds is on the extraction path: "
                                      "resnet18, depth 4, decoder (256,128,64,32), batchnorm, no attention, 3 -> 3, no activation")
        if encoder_weights is not None:
            # smp downloads the ImageNet ResNet-18 here; there is no network on the extraction boxes and every
            # reference caller loads an `fa_model_state_dict` checkpoint right afterwards (dali_extraction.py:122)
            warnings.warn("UnetPlusPlus: encoder_weights=%r is not downloaded; load a checkpoint (load_fa_model("
                          "saved_model_file=...)) - the encoder starts from torch's default init" % (encoder_weights,))
        self.encoder = _ResNet18Encoder()
        self.decoder = _Decoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(32, classes, 3, padding=1), nn.Identity(), nn.Identity())

    def forward(self, x):
        ex = self._exec(x)
        n, _, h, w = x.shape
        if h % 16 or w % 16:
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width divisible by 16.")
        with torch.cuda.device(x.device):
            from tedspad_b200 import ops
            x0 = ex.input_buffer(n, h, w)
            ops.nchw_to_cl(x, x0)
            clip = ex.bufs.get("out_clip", n, 1, h, w, 4)
            out = ex.bufs.raw("frames_out", (n, 3, h, w), torch.float32)
            self._graphed(ex, ("forward", n, h, w), lambda: ex.run(x0, clip, T=1, frames_out=out))
            return out.clone()

    def anonymize_into(self, x0, enc_in, T=16):
        """Fused path of the extraction driver: frames already in `ex.input_buffer` layout -> anonymized planes
        written straight into the encoder input through the raw-reshape glue (dali_extraction.py:171-173)."""
        return self._exec(x0.buf).run(x0, enc_in, T=T)
