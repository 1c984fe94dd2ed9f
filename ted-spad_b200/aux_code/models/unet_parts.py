"""Parameter containers of the anonymizer UNet.  They reproduce the reference's module tree
(aux_code/models/unet_parts.py:8-77 in UCF-CRCV/TeD-SPAD) so that `state_dict()` keys and shapes
are identical (`inc.double_conv.0.weight`, `down1.maxpool_conv.1.double_conv.4.running_var`, ...);
the arithmetic is done by tedspad_b200.engine.UNetExecutor, not by these modules."""
import torch.nn as nn


def _conv_bn_relu(cin, cout):
    return [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        self.double_conv = nn.Sequential(*_conv_bn_relu(in_channels, mid), *_conv_bn_relu(mid, out_channels))


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if not bilinear:
            raise NotImplementedError("only the bilinear UNet the reference builds (unet_model.py:6) is supported")
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)
