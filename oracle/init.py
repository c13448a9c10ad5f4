"""Oracle (test infrastructure): random-init state dict with the reference's RNG consumption.

Follows reference models/resnet_language.py ResNet.__init__ :101-141 and _make_layer :143-167:
construction order per stage is downsample(Conv2d 1x1, BatchNorm2d) first, then the block's conv1/bn1/conv2/bn2/
conv3/bn3 (each nn.Conv2d draws its default kaiming-uniform init); after all stages every Conv2d is re-drawn with
kaiming_normal_(fan_out, leaky_relu) in ``self.modules()`` order (conv1, conv2, conv3, downsample.0 per block), BN
affine = (1, 0); the classifier nn.Linear(640, n_cls) is created last with its default init.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .backbone import block_plan


def init_state_dict(seed, n_cls=60, model='resnet18', linear_bias=False):
    torch.manual_seed(seed)
    plan = block_plan(model, True)
    mods = OrderedDict()
    for blk in plan:
        p, cin, cout = blk['prefix'], blk['cin'], blk['cout']
        if blk['downsample']:
            ds = (nn.Conv2d(cin, cout, kernel_size=1, stride=1, bias=False), nn.BatchNorm2d(cout))
        convs = []
        for i, ci in enumerate((cin, cout, cout)):
            convs.append((nn.Conv2d(ci, cout, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(cout)))
        for i, (c, b) in enumerate(convs):
            mods[p + '.conv%d' % (i + 1)] = c
            mods[p + '.bn%d' % (i + 1)] = b
        if blk['downsample']:
            mods[p + '.downsample.0'] = ds[0]
            mods[p + '.downsample.1'] = ds[1]
    for name, m in mods.items():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='leaky_relu')
    mods['classifier'] = nn.Linear(640, n_cls, bias=linear_bias)
    sd = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[name + '.' + k] = v.detach().clone()
    return sd
