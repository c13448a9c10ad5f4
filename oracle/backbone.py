"""Oracle (test infrastructure): functional restatement of the RFS ResNet-12/18 backbone on a plain state dict.

Follows reference models/resnet_language.py:
  ResNet.__init__ / _make_layer  :101-167   (stage plan: 64/160/320/640 planes, blocks [1,1,2,2] for resnet18)
  ResNet.forward                 :170-192
  BasicBlock.forward             :268-301
  DropBlock.forward / mask       :311-357
PyTorch CPU fp32 throughout; BatchNorm train/eval semantics are torch.nn.functional.batch_norm's.
"""
import torch
import torch.nn.functional as F

PLANES = (64, 160, 320, 640)
N_BLOCKS = {'resnet12': (1, 1, 1, 1), 'resnet18': (1, 1, 2, 2)}
DROP_RATE = 0.1            # models/util.py:15-18 passes drop_rate=0.1, dropblock_size=5
DROPBLOCK_SIZE = 5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SLOPE = 0.1


def block_plan(model='resnet18', no_dropblock=True):
    """[(prefix, cin, cout, pool_stride, has_downsample, drop_block, block_size)] in forward order.

    resnet_language.py:112-122,143-167: every stage's FIRST block carries the 1x1 downsample and MaxPool2d(2).
    Quirk kept from the reference (:155): for multi-block stages the first block is built as
    ``block(inplanes, planes, stride, downsample, drop_rate, self.use_se)`` so ``use_se`` (False) lands in the
    ``drop_block`` slot -> plain dropout; the LAST block of layer3/layer4 gets drop_block=True (:158-160).
    For single-block stages (:153) layer3/layer4's only block gets drop_block=True."""
    plan = []
    cin = 3
    block_size = 1 if no_dropblock else DROPBLOCK_SIZE
    for li, (planes, nb) in enumerate(zip(PLANES, N_BLOCKS[model])):
        stage_drop_block = li >= 2  # layer3 / layer4 are built with drop_block=True (:119-122)
        for bi in range(nb):
            first = bi == 0
            last = bi == nb - 1
            if nb == 1:
                db = stage_drop_block
            else:
                db = stage_drop_block and last and not first
            plan.append(dict(prefix='layer%d.%d' % (li + 1, bi), cin=cin, cout=planes, pool=2 if first else 1,
                             downsample=first, drop_block=db, block_size=block_size if db else 1))
            cin = planes
    return plan


def new_counters(plan):
    """BasicBlock.num_batches_tracked: a python int per block, +1 on EVERY forward (:260,269)."""
    return {b['prefix']: 0 for b in plan}


def _bn(sd, name, x, train):
    rm, rv = sd[name + '.running_mean'], sd[name + '.running_var']
    if train:
        sd[name + '.num_batches_tracked'] += 1      # nn.BatchNorm2d.forward bookkeeping in train mode
    return F.batch_norm(x, rm, rv, sd[name + '.weight'], sd[name + '.bias'], training=train, momentum=BN_MOMENTUM,
                        eps=BN_EPS)


def dropblock_mask(mask, block_size):
    """DropBlock._compute_block_mask (:327-357): each sampled seed zeroes a block_size^2 block whose top-left corner is
    the seed position in the padded map."""
    left, right = int((block_size - 1) / 2), int(block_size / 2)
    padded = F.pad(mask, (left, right, left, right))
    Hm, Wm = mask.shape[2], mask.shape[3]
    if mask.any():
        for i in range(block_size):
            for j in range(block_size):
                region = padded[:, :, i:i + Hm, j:j + Wm]
                padded[:, :, i:i + Hm, j:j + Wm] = torch.maximum(region, mask)
    return 1 - padded


def dropblock_gamma(nbt, feat_size, block_size):
    """BasicBlock.forward :295-296."""
    keep_rate = max(1.0 - DROP_RATE / (20 * 2000) * nbt, 1.0 - DROP_RATE)
    return (1 - keep_rate) / block_size ** 2 * feat_size ** 2 / (feat_size - block_size + 1) ** 2


def block_forward(sd, blk, x, train, counters):
    p = blk['prefix']
    counters[p] += 1
    out = F.conv2d(x, sd[p + '.conv1.weight'], padding=1)
    out = F.leaky_relu(_bn(sd, p + '.bn1', out, train), SLOPE)
    out = F.conv2d(out, sd[p + '.conv2.weight'], padding=1)
    out = F.leaky_relu(_bn(sd, p + '.bn2', out, train), SLOPE)
    out = F.conv2d(out, sd[p + '.conv3.weight'], padding=1)
    out = _bn(sd, p + '.bn3', out, train)
    if blk['downsample']:
        residual = _bn(sd, p + '.downsample.1', F.conv2d(x, sd[p + '.downsample.0.weight']), train)
    else:
        residual = x
    out = F.leaky_relu(out + residual, SLOPE)
    out = F.max_pool2d(out, blk['pool'])
    # drop_rate > 0 always (0.1)
    if blk['drop_block']:
        if train:
            bs = blk['block_size']
            gamma = dropblock_gamma(counters[p], out.shape[2], bs)
            B, C, H, W = out.shape
            seeds = torch.distributions.Bernoulli(gamma).sample((B, C, H - (bs - 1), W - (bs - 1)))
            bm = dropblock_mask(seeds, bs)
            out = bm * out * (bm.numel() / bm.sum())
    else:
        out = F.dropout(out, p=DROP_RATE, training=train)
    return out


def features(sd, plan, x, train, counters):
    """ResNet.forward up to `feat` (:170-181): four stages, AdaptiveAvgPool2d(1), flatten -> [B, 640]."""
    for blk in plan:
        x = block_forward(sd, blk, x, train, counters)
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


def forward(sd, plan, x, train, counters):
    """-> logits = feat @ classifier.weight^T (+ bias) (:182-187)."""
    return F.linear(features(sd, plan, x, train, counters), sd['classifier.weight'], sd.get('classifier.bias'))
