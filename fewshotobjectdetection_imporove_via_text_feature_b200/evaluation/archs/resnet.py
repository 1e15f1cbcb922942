"""ImageNet ResNet-101 used by the Prototypical Calibration Block as a frozen feature extractor.

Interface mirror of defrcn/evaluation/archs/resnet.py (`resnet101()`; `model(x) -> (logits, layer4 feature)`,
`model.fc`) with torchvision's parameter names (conv1, bn1, layer{1..4}.{i}.conv{1,2,3} / bn{1,2,3} / downsample.{0,1},
fc), so the checkpoint cfg.TEST.PCB_MODELPATH points at loads unchanged.  Outside the hot path:
plain torch modules — the convolutions run wherever torch runs them."""
import torch
from torch import nn

_STAGES = ((64, 3, 1), (128, 4, 2), (256, 23, 2), (512, 3, 2))     # (bottleneck width, blocks, stride of the first block)


class _Block(nn.Module):
    """1x1 reduce -> 3x3 (carries the stride) -> 1x1 expand (x4), identity or projected skip, ReLU after the sum."""

    def __init__(self, cin, width, stride):
        super().__init__()
        cout = 4 * width
        self.conv1, self.bn1 = nn.Conv2d(cin, width, 1, bias=False), nn.BatchNorm2d(width)
        self.conv2, self.bn2 = nn.Conv2d(width, width, 3, stride, 1, bias=False), nn.BatchNorm2d(width)
        self.conv3, self.bn3 = nn.Conv2d(width, cout, 1, bias=False), nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = torch.relu(self.bn1(self.conv1(x)))
        y = torch.relu(self.bn2(self.conv2(y)))
        return torch.relu(self.bn3(self.conv3(y)) + skip)


class ResNet101(nn.Module):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for i, (width, blocks, stride) in enumerate(_STAGES, 1):
            layers = []
            for b in range(blocks):
                layers.append(_Block(cin, width, stride if b == 0 else 1))
                cin = 4 * width
            setattr(self, "layer%d" % i, nn.Sequential(*layers))
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = self.maxpool(torch.relu(self.bn1(self.conv1(x))))
        feature = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(feature.mean(dim=(2, 3))), feature


def resnet101(**kwargs):
    return ResNet101(**kwargs)
