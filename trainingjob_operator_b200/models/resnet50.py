"""ResNet-50 (bf16 autocast, channels_last) -- BASELINE.json config "ResNet-50 DDP bf16 elastic min=2 max=8".

The reference has no model code; this is one of the launched workers' benchmark networks: the standard
[3, 4, 6, 3] bottleneck stack (stride on the 3x3 convolution, zero-initialised last BatchNorm of every block), written
out here rather than imported.  Convolutions and BatchNorm are cuDNN calls (plain library ops); what this repo adds around
them is the training step -- forward + backward replayed as ONE CUDA graph, the flat-buffer gradient reduction, the fused
AdamW sweep and the elastic state hand-off (``parallel.flat_ddp``, ``runtime.worker.TorchAdapter``).  Synthetic
3x224x224 images, random-init weights."""
from __future__ import annotations

import torch
import torch.nn as nn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin: int, mid: int, stride: int):
        super().__init__()
        cout = mid * self.expansion
        self.conv1 = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.conv3 = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        nn.init.zeros_(self.bn3.weight)            # every residual branch starts as the identity
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.down is None else self.down(x)
        y = torch.relu(self.bn1(self.conv1(x)))
        y = torch.relu(self.bn2(self.conv2(y)))
        return torch.relu(self.bn3(self.conv3(y)) + idt)


class ResNet50(nn.Module):
    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool2d(3, 2, 1))
        blocks, cin = [], 64
        for mid, n, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for i in range(n):
                blocks.append(Bottleneck(cin, mid, stride if i == 0 else 1))
                cin = mid * Bottleneck.expansion
        self.blocks = nn.Sequential(*blocks)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        return self.fc(torch.flatten(self.pool(self.blocks(self.stem(x))), 1))


def build_resnet50(num_classes: int = 1000) -> nn.Module:
    return ResNet50(num_classes)
