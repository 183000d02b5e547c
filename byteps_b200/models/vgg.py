"""VGG-16/19 (the reference's bandwidth-heavy CNN: 138 M parameters, most of
them in three fully connected layers)."""
import torch
import torch.nn as nn

_CFG = {
    16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    def __init__(self, depth=16, num_classes=1000, batch_norm=False, dropout=0.5):
        super().__init__()
        layers, c = [], 3
        for v in _CFG[depth]:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers.append(nn.Conv2d(c, v, 3, padding=1))
                if batch_norm:
                    layers.append(nn.BatchNorm2d(v))
                layers.append(nn.ReLU(inplace=True))
                c = v
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential(
            nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(dropout),
            nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(dropout), nn.Linear(4096, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.avgpool(self.features(x)), 1))


def vgg16(**kw):
    return VGG(16, **kw)


def vgg19(**kw):
    return VGG(19, **kw)
