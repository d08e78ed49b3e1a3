"""2D backbone: ResNet-50 trunk + three stride-2 deconvolutions + 1x1 head.

Out of scope for the B200 path (BASELINE.json north_star: "after the pose_resnet
backbone"); it stays stock PyTorch and only provides the stage's input.  Module
and parameter names follow the reference (network/pose_resnet.py:135-246) so
its checkpoints load: conv1, bn1, layer1..4.{i}.{conv1..3,bn1..3,downsample},
deconv_layers.{0..8}, final_layer.
"""
import torch
import torch.nn as nn

BN_MOMENTUM = 0.1


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + idt)


class PoseResNet(nn.Module):
    def __init__(self, blocks=(3, 4, 6, 3), num_outputs=16):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = self._stage(64, blocks[0], 1)
        self.layer2 = self._stage(128, blocks[1], 2)
        self.layer3 = self._stage(256, blocks[2], 2)
        self.layer4 = self._stage(512, blocks[3], 2)
        up = []
        for _ in range(3):
            up += [nn.ConvTranspose2d(self.inplanes, 256, 4, stride=2, padding=1, output_padding=0, bias=False),
                   nn.BatchNorm2d(256, momentum=BN_MOMENTUM), nn.ReLU(inplace=True)]
            self.inplanes = 256
        self.deconv_layers = nn.Sequential(*up)
        self.final_layer = nn.Conv2d(256, num_outputs, 1)

    def _stage(self, planes, n, stride):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(planes * 4, momentum=BN_MOMENTUM))
        layers = [Bottleneck(self.inplanes, planes, stride, down)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes) for _ in range(1, n)]
        return nn.Sequential(*layers)

    def forward_before_last_deconv(self, x):
        """Everything up to the input of the last deconvolution stage: (B,256,32,32).  The hand-off kernel
        (csrc/handoff.cu) fuses that stage with the volumetric stage's 1x1 conv."""
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.deconv_layers[:6](x)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        features = self.deconv_layers(x)
        heatmaps = self.final_layer(features)[:, :15]
        return heatmaps, features


def get_pose_net(model_path=None, state_dict=None):
    """network/pose_resnet.py:313-332: optional weights, `module.` prefix stripped."""
    model = PoseResNet()
    if state_dict is None and model_path is not None:
        state_dict = torch.load(model_path)
    if state_dict is not None:
        if next(iter(state_dict)).startswith("module"):
            state_dict = {k[7:]: v for k, v in state_dict.items()}
        merged = model.state_dict()
        merged.update(state_dict)
        model.load_state_dict(merged)
    return model
