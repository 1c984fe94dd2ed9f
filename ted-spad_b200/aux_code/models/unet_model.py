"""Anonymizer UNet (`load_fa_model(arch='unet')`), drop-in for the reference's
aux_code/models/unet_model.py:6-37: same constructor, same parameter names, same
`forward(x[N,3,H,W] in [0,1]) -> [N,3,H,W] in (0,1)`, computed by the tcgen05 executor."""
import torch

from aux_code._base import CLTensor, CudaModule
from aux_code.models.unet_parts import DoubleConv, Down, OutConv, Up
from tedspad_b200.engine import UNetExecutor


class UNet(CudaModule):
    executor_cls = UNetExecutor

    def __init__(self, n_channels, n_classes, bilinear=True):
        super().__init__()
        if n_channels != 3 or n_classes != 3 or not bilinear:
            raise NotImplementedError("UNet(n_channels=3, n_classes=3, bilinear=True) is the only configuration on the "
                                      "extraction path (model_loaders.py:32)")
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1, self.down2, self.down3 = Down(64, 128), Down(128, 256), Down(256, 512)
        self.down4 = Down(512, 512)
        self.up1, self.up2, self.up3, self.up4 = Up(1024, 256), Up(512, 128), Up(256, 64), Up(128, 64)
        self.outc = OutConv(64, n_classes)
        self.sigm = torch.nn.Sigmoid()

    def forward(self, x):
        ex = self._exec(x)
        n, _, h, w = x.shape
        with torch.cuda.device(x.device):
            x0 = ex.input_buffer(n, h, w)
            self._fill(ex, x, x0)
            nhwc = ex.bufs.get("out_nhwc", n, 1, h, w, 8)
            out = ex.bufs.raw("frames_out", (n, 3, h, w), torch.float32)
            self._graphed(ex, ("forward", n, h, w), lambda: ex.run(x0, nhwc, T=1, frames_out=out))
            return out.clone()

    @staticmethod
    def _fill(ex, x, x0):
        from tedspad_b200 import ops
        ops.nchw_to_cl(x, x0)

    def anonymize_into(self, x0, enc_in, T=16):
        """Fused path used by the extraction driver: frames already in `ex.input_buffer` layout ->
        anonymized planes scattered straight into the encoder input (dali_extraction.py:171-173)."""
        ex = self._exec(x0.buf)
        return ex.run(x0, enc_in, T=T)

