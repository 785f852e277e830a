"""MobileNetV2 feature pyramid for SMOW_Net_LW (cuDNN; not hot path).

Key-compatible with the reference's vendored copy (models/SMOW_Net_LW.py:482-589), i.e. with
torchvision's ``mobilenet_v2`` checkpoints: ``features.N`` blocks, five taps after blocks
1, 3, 6, 13 and 17 (16/24/32/96/320 channels at strides 2/4/8/16/32).  ``features.18`` is
executed although its output is unused, exactly like the reference (:573-579), so that its
BatchNorm running statistics evolve identically during training.
"""
import torch.nn as nn

WEIGHTS_URL = "https://download.pytorch.org/models/mobilenet_v2-b0353104.pth"
# expansion, channels, repeats, stride  (dilation is 1 everywhere in this configuration)
STAGES = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
TAPS = (1, 3, 6, 13, 17)


def conv_bn_relu6(cin, cout, kernel=3, stride=1, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel, stride, (kernel - 1) // 2, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        hidden = int(round(cin * expand))
        self.use_res_connect = stride == 1 and cin == cout
        seq = [] if expand == 1 else [conv_bn_relu6(cin, hidden, kernel=1)]
        seq += [conv_bn_relu6(hidden, hidden, stride=stride, groups=hidden),
                nn.Conv2d(hidden, cout, 1, 1, 0, bias=False), nn.BatchNorm2d(cout)]
        self.conv = nn.Sequential(*seq)

    def forward(self, x):
        y = self.conv(x)
        return x + y if self.use_res_connect else y


class MobileNetV2Pyramid(nn.Module):
    def __init__(self):
        super().__init__()
        self.last_channel = 1280
        layers = [conv_bn_relu6(3, 32, stride=2)]
        cin = 32
        for expand, cout, repeats, stride in STAGES:
            for i in range(repeats):
                layers.append(InvertedResidual(cin, cout, stride if i == 0 else 1, expand))
                cin = cout
        layers.append(conv_bn_relu6(cin, self.last_channel, kernel=1))
        self.features = nn.Sequential(*layers)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        taps = []
        for idx, block in enumerate(self.features):
            x = block(x)
            if idx in TAPS:
                taps.append(x)
        return taps


def mobilenet_v2(pretrained=True, progress=True):
    model = MobileNetV2Pyramid()
    if pretrained:
        from torch.hub import load_state_dict_from_url
        model.load_state_dict(load_state_dict_from_url(WEIGHTS_URL, progress=progress), strict=False)
    return model
