"""Mirror of the reference's live saliency ("localization") network, saliency_network.py:302-333.

Stock PyTorch (cuDNN convolutions): the north star keeps the networks stock; this file exists so that the drop-in
package is importable under the reference's module names and checkpoints (`saliency_epoch_*.pth`) load unchanged:
parameter names fov_expand_1/2, fov_squeeze_1, norm1-3 are the reference's.  `SynchronizedBatchNorm2d` of the
reference degenerates to plain batch norm under DDP (lib/nn/modules/batchnorm.py:58-61), hence nn.BatchNorm2d.
"""
import torch
import torch.nn as nn


class FovSimModule(nn.Module):
    def __init__(self, cfg, in_channels=5, out_channels=3):
        super().__init__()
        self.cfg = cfg
        self.fov_expand_1 = nn.Conv2d(in_channels, 8 * out_channels, kernel_size=3, padding=1, bias=False)
        self.fov_expand_2 = nn.Conv2d(8 * out_channels, 8 * out_channels, kernel_size=3, padding=1, bias=False)
        self.fov_squeeze_1 = nn.Conv2d(8 * out_channels, out_channels, kernel_size=3, padding=1, bias=False)
        self.norm1 = nn.BatchNorm2d(8 * out_channels, momentum=0.1)
        self.norm2 = nn.BatchNorm2d(8 * out_channels, momentum=0.1)
        self.norm3 = nn.BatchNorm2d(out_channels, momentum=0.1)
        self.act = nn.ReLU6(inplace=False)

    def forward(self, x, reset_grad=True, train_mode=True):
        layer1 = self.act(self.norm1(self.fov_expand_1(x)))
        layer2 = self.act(self.norm2(self.fov_expand_2(layer1)))
        return self.norm3(self.fov_squeeze_1(layer2))


def fov_simple(cfg, pretrained=False, in_channels=5, out_channels=24):
    model = FovSimModule(cfg, in_channels=in_channels, out_channels=out_channels)
    if pretrained:
        path = "./pretrained/foveater_cityscape_soft_e100.pth"
        model.load_state_dict(torch.load(path, map_location=lambda storage, loc: storage), strict=False)
    return model
