"""VGG-16 (configuration D, Simonyan & Zisserman 2014): 13 conv3x3 + 3 FC, 138.4 M parameters."""
import torch
import torch.nn as nn

_CFG_D = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


class VGG(nn.Module):
    def __init__(self, cfg=_CFG_D, num_classes: int = 1000, dropout: float = 0.5):
        super().__init__()
        self.fuse_epilogues = True  # use the fused NHWC conv-block epilogues on CUDA f16/bf16 channels_last inputs
        layers, c_in = [], 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(c_in, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                c_in = v
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential(
            nn.Linear(512 * 7 * 7, 4096), nn.ReLU(True), nn.Dropout(p=dropout),
            nn.Linear(4096, 4096), nn.ReLU(True), nn.Dropout(p=dropout),
            nn.Linear(4096, num_classes),
        )
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def _features_fused(self, x):
        """conv → (bias+ReLU[+maxpool]) blocks through the fused NHWC epilogue kernels; module structure / state_dict stay
        the standard ``features.N`` Sequential."""
        from ..ops.nhwc import conv_bias_relu

        mods = list(self.features)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                pool = i + 2 < len(mods) and isinstance(mods[i + 2], nn.MaxPool2d) and mods[i + 2].kernel_size == 2 and mods[i + 2].stride == 2
                if x.is_contiguous(memory_format=torch.channels_last) and m.out_channels % 8 == 0:
                    x = conv_bias_relu(x, m, pool)
                    i += 3 if pool else 2
                    continue
            x = m(x)
            i += 1
        return x

    def forward(self, x):
        if self.fuse_epilogues and x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.dim() == 4:
            x = self._features_fused(x.contiguous(memory_format=torch.channels_last))
        else:
            x = self.features(x)
        x = self.avgpool(x)
        x = torch.flatten(x, 1)
        return self.classifier(x)


def vgg16(num_classes: int = 1000, **kw) -> VGG:
    return VGG(_CFG_D, num_classes, **kw)
