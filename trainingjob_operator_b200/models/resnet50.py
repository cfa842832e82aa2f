"""ResNet-50 (bf16, channels_last) -- BASELINE.json config "ResNet-50 DDP bf16 elastic min=2 max=8".

The reference has no model code.  The network definition is torchvision's
(convolutions are cuDNN library calls -- allowed as plain library ops); the data-parallel gradient
reduction, the flat-buffer fused AdamW/SGD sweep and the elastic state hand-off around it are this
repo's (``parallel.flat_ddp``).  Synthetic 3x224x224 images, random-init weights."""
from __future__ import annotations

import torch


def build_resnet50(num_classes: int = 1000) -> torch.nn.Module:
    try:
        from torchvision.models import resnet50

        return resnet50(weights=None, num_classes=num_classes)
    except Exception:  # noqa: BLE001 - torchvision missing: small bottleneck stack with the same interface
        import torch.nn as nn

        class Bottleneck(nn.Module):
            def __init__(self, cin, mid, cout, stride):
                super().__init__()
                self.c1 = nn.Conv2d(cin, mid, 1, bias=False); self.b1 = nn.BatchNorm2d(mid)
                self.c2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False); self.b2 = nn.BatchNorm2d(mid)
                self.c3 = nn.Conv2d(mid, cout, 1, bias=False); self.b3 = nn.BatchNorm2d(cout)
                self.down = None
                if stride != 1 or cin != cout:
                    self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

            def forward(self, x):
                idt = x if self.down is None else self.down(x)
                y = torch.relu(self.b1(self.c1(x)))
                y = torch.relu(self.b2(self.c2(y)))
                return torch.relu(self.b3(self.c3(y)) + idt)

        layers, cin = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(3, 2, 1)], 64
        for mid, n, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
            for i in range(n):
                layers.append(Bottleneck(cin, mid, mid * 4, stride if i == 0 else 1))
                cin = mid * 4
        layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(cin, num_classes)]
        return nn.Sequential(*layers)
