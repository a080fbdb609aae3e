"""Tiny stand-ins for the (out-of-scope, stock PyTorch) encoder / decoder, shared by the golden generator and the
GPU module test so that both sides run identical weights."""
import torch
import torch.nn as nn


class TinyEncoder(nn.Module):
    def __init__(self, cout=8):
        super().__init__()
        self.conv = nn.Conv2d(3, cout, 3, padding=1)

    def forward(self, x, return_feature_maps=False):
        return [torch.relu(self.conv(x))]


class TinyDecoder(nn.Module):
    def __init__(self, cin=8, num_class=51):
        super().__init__()
        self.conv = nn.Conv2d(cin, num_class, 1)

    def forward(self, feats, segSize=None):
        return self.conv(feats[-1])


def synthetic_batch(B, H, W, seed):
    """image ~ U[0,1), gaze ~ U[0,0.98)^2, label = disc around the gaze, class ~ randint(0,50)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, H, W, generator=g)
    gaze = torch.rand(B, 2, generator=g) * 0.98
    rad = (0.03 + 0.22 * torch.rand(B, generator=g)) * H
    ii = torch.arange(H, dtype=torch.float32)[None, :, None]
    jj = torch.arange(W, dtype=torch.float32)[None, None, :]
    d2 = (ii - gaze[:, 0, None, None] * (H - 1)) ** 2 + (jj - gaze[:, 1, None, None] * (W - 1)) ** 2
    y = (d2 <= rad[:, None, None] ** 2).float().unsqueeze(1)
    cls = torch.randint(0, 50, (B, 1), generator=g)
    return {"img_data": x, "seg_label": y, "focus_point": gaze, "cls_label": cls}
