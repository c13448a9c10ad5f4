"""Drop-in for reference models/resnet_language.py (ResNet :101-240, BasicBlock :243-301, LangPuller :20-97,
LinearMap :12-18, resnet12/resnet18 :409-419): same class / method names, state-dict keys and RNG consumption;
all arithmetic runs in the sm_100a kernels of libsrb200.so (there is no CPU path - tensors must be on a B200).
"""
import os

import torch
import torch.nn as nn

from srb200 import autograd as sra
from srb200 import ops
from srb200 import rng
from srb200.backbone import BackboneEngine
from .util import get_embeds


class LinearMap(nn.Module):
    """nn.Linear(indim, outdim) under the attribute name `map` (state-dict keys map.weight / map.bias)."""

    def __init__(self, indim, outdim):
        super(LinearMap, self).__init__()
        self.map = nn.Linear(indim, outdim)

    def forward(self, x):
        return sra.linear(x, self.map.weight, self.map.bias)


class LangPuller(nn.Module):
    """Targets ("pullers") for the newest classifier rows: label-embedding softmax mix of base weights, a learned
    linear map of label embeddings, or the projection onto span(base weights)."""

    def __init__(self, opt, vocab_base, vocab_novel):
        super(LangPuller, self).__init__()
        if opt.use_synonyms:
            raise NotImplementedError("use_synonyms needs {dataset}_dim{dim}_base_synonyms.pickle, which the reference "
                                      "does not ship")
        self.mapping_model = None
        self.opt = opt
        self.vocab_base = vocab_base
        self.vocab_novel = vocab_novel
        self.temp = opt.temperature
        self._path = os.path.join(opt.word_embed_path, "{0}_dim{1}.pickle".format(opt.dataset, opt.word_embed_size))
        self.novel_embeds = self._load(vocab_novel)
        self.base_embeds = self._load(vocab_base)
        self._factor = None      # (key, qt, q_rows, identity) cache: W0 is constant for a whole run

    def _load(self, vocab):
        desc = getattr(self.opt, 'description_embed_path', None)
        if desc is not None:
            # BASELINE config 3 (ii): label-description embeddings (description_embeds/*.pickle: dict label -> Tensor[768]).
            # The reference ships the pickles but no consumer; they enter through the same two puller modes.
            import pickle
            with open(desc, "rb") as f:
                table = pickle.load(f)
            return torch.stack([torch.as_tensor(table[name]).float() for name in vocab], 0).cuda().contiguous()
        e = get_embeds(self._path, vocab).float().cuda()
        if self.opt.glove:       # the first 300 dims of the saved vectors are GloVe
            e = e[:, :300].contiguous()
        return e

    def update_novel_embeds(self, vocab_novel):
        self.vocab_novel = vocab_novel
        self.novel_embeds = self._load(vocab_novel)

    def create_pulling_mapping(self, state_dict, base_weight_size=640):
        if rng.private():
            # (a run on a private generator: consume the draws LinearMap's nn.Linear init makes, then build the module
            # without touching the process-global generator's stream semantics)
            rng.linear_init(base_weight_size, self.novel_embeds.size(1), True)
            saved = torch.get_rng_state()
            self.mapping_model = LinearMap(self.novel_embeds.size(1), base_weight_size)
            torch.set_rng_state(saved)
        else:
            self.mapping_model = LinearMap(self.novel_embeds.size(1), base_weight_size)
        self.mapping_model.load_state_dict(state_dict)
        self.mapping_model = self.mapping_model.cuda()

    def forward(self, base_weight, mask=False):
        if self.mapping_model is None:
            return ops.semantic_pullers(self.novel_embeds, self.base_embeds, base_weight.detach().contiguous(),
                                        float(self.temp), bool(mask))
        with torch.no_grad():
            return self.mapping_model(self.novel_embeds)

    def loss1(self, pull, inspired, weights):
        return sra.scaled_sqdist(float(pull), inspired, weights)

    def factor(self, base_weight):
        key = (base_weight.data_ptr(), base_weight._version, tuple(base_weight.shape))
        if self._factor is None or self._factor[0] != key:
            qt, q, ident = ops.subspace_factor(base_weight.detach().contiguous())
            self._factor = (key, qt, q, ident)
        return self._factor[1:]

    def get_projected_weight(self, base_weight, weights):
        qt, q, ident = self.factor(base_weight)
        return sra.project_rows(weights, qt, q, ident)


class BasicBlock(nn.Module):
    """Parameter container with the reference's attribute names; the arithmetic is sequenced by BackboneEngine."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, drop_rate=0.0, drop_block=False, block_size=1,
                 use_se=False):
        super(BasicBlock, self).__init__()
        if use_se:
            raise NotImplementedError("SE blocks are not reachable from model_pool")
        self.conv1 = conv3x3(inplanes, planes)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = conv3x3(planes, planes)
        self.bn3 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.drop_rate = drop_rate
        self.num_batches_tracked = 0      # python int, +1 on EVERY forward of the block
        self.drop_block = drop_block
        self.block_size = block_size

    def forward(self, x):
        raise RuntimeError("BasicBlock is driven by ResNet.forward (fused sm_100a kernels); it has no stand-alone path")


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class ResNet(nn.Module):

    def __init__(self, block, n_blocks, keep_prob=1.0, avg_pool=False, drop_rate=0.0, dropblock_size=5, num_classes=-1,
                 use_se=False, vocab=None, opt=None):
        if vocab is not None:
            assert opt is not None
            raise NotImplementedError("language classifiers (vocab != None) are not on the incremental-session path")
        super(ResNet, self).__init__()
        self.inplanes = 3
        self.use_se = use_se
        self.layer1 = self._make_layer(block, n_blocks[0], 64, stride=2, drop_rate=drop_rate)
        self.layer2 = self._make_layer(block, n_blocks[1], 160, stride=2, drop_rate=drop_rate)
        if opt.no_dropblock:
            dropblock_size = 1
        self.layer3 = self._make_layer(block, n_blocks[2], 320, stride=2, drop_rate=drop_rate, drop_block=True,
                                       block_size=dropblock_size)
        self.layer4 = self._make_layer(block, n_blocks[3], 640, stride=2, drop_rate=drop_rate, drop_block=True,
                                       block_size=dropblock_size)
        if not avg_pool:
            raise NotImplementedError("avg_pool=False is never used by create_model")
        self.keep_prob = keep_prob
        self.keep_avg_pool = avg_pool
        self.drop_rate = drop_rate
        self.vocab = vocab
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='leaky_relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.num_classes = num_classes
        if self.num_classes > 0:
            self.classifier = nn.Linear(640, self.num_classes, bias=opt.linear_bias)
        self._engine = None

    def _make_layer(self, block, n_block, planes, stride=1, drop_rate=0.0, drop_block=False, block_size=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=1, bias=False),
                nn.BatchNorm2d(planes * block.expansion),
            )
        layers = []
        if n_block == 1:
            layers.append(block(self.inplanes, planes, stride, downsample, drop_rate, drop_block, block_size, self.use_se))
        else:
            # reference quirk kept (resnet_language.py:155): use_se lands in the drop_block slot -> plain dropout
            layers.append(block(self.inplanes, planes, stride, downsample, drop_rate, self.use_se))
        self.inplanes = planes * block.expansion
        for i in range(1, n_block):
            if i == n_block - 1:
                layers.append(block(self.inplanes, planes, drop_rate=drop_rate, drop_block=drop_block,
                                    block_size=block_size, use_se=self.use_se))
            else:
                layers.append(block(self.inplanes, planes, drop_rate=drop_rate, use_se=self.use_se))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ engine plumbing
    def _blocks(self):
        out = []
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            for bi, m in enumerate(layer):
                out.append(dict(prefix='layer%d.%d' % (li + 1, bi), mod=m, cin=m.conv1.weight.shape[1],
                                cout=m.conv1.weight.shape[0], pool=m.stride, downsample=m.downsample is not None,
                                drop_block=bool(m.drop_block), block_size=m.block_size, drop_rate=float(m.drop_rate)))
        return out

    def engine(self):
        if self._engine is None:
            self._engine = BackboneEngine(self._blocks(), precision=getattr(self, '_conv_precision', None))
        return self._engine

    def set_conv_precision(self, precision):
        """'bf16' (one tensor-core pass, the throughput tier) or 'bf16x3' (error-compensated operand pairs, three passes
        into one fp32 accumulator: the parity tier against the reference's fp32 convolutions)."""
        self._conv_precision = precision
        self.engine().set_precision(precision)

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_engine' else copy.deepcopy(v, memo)
        return new

    def block_counters(self):
        return {b['prefix']: b['mod'].num_batches_tracked for b in self._blocks()}

    def advance_block_counters(self, n):
        """Account for `n` backbone forwards that the feature cache made unnecessary (DropBlock's gamma schedule reads
        BasicBlock.num_batches_tracked, which the reference bumps on every forward, eval included)."""
        for b in self._blocks():
            b['mod'].num_batches_tracked += n

    def features(self, x, taps=None):
        """[B,3,84,84] fp32 CUDA -> pooled 640-d features, in the module's current mode (train: batch-stat BN with
        running-stat update + dropout / DropBlock; eval: folded BN)."""
        if not x.is_cuda:
            raise RuntimeError("srb200: ResNet.forward needs CUDA tensors on a B200 (no CPU path)")
        raw_u8 = x.dtype == torch.uint8 and x.dim() == 4 and tuple(x.shape[1:]) == (84, 84, 3)   # image store, NHWC
        if not raw_u8 and (x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != 84 or x.shape[3] != 84):
            raise RuntimeError("srb200: expected fp32 [B,3,84,84] (or uint8 [B,84,84,3]) input, got %s %s" %
                               (x.dtype, tuple(x.shape)))
        self.advance_block_counters(1)
        eng = self.engine()
        if self.training:
            head = {id(p) for p in self.classifier.parameters()} if self.num_classes > 0 else ()
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters() if id(p) not in head):
                raise NotImplementedError("backbone training is outside the incremental-session path "
                                          "(use --freeze_backbone_at 1)")
            return eng.train_features(x, self.block_counters())
        return eng.eval_features(x, taps)

    def forward(self, x, is_feat=False, get_alphas=False):
        taps = [] if is_feat else None
        feat = self.features(x, taps)
        out = feat
        if self.num_classes > 0:
            out = sra.linear(feat, self.classifier.weight, self.classifier.bias)
        if is_feat:
            if self.training:
                raise NotImplementedError("is_feat=True is only available in eval mode")
            maps = [(t[0].float() + t[1].float() if t.dim() == 5 else t.float()).permute(0, 3, 1, 2) for t in taps]
            return maps + [feat], out
        return out

    # ------------------------------------------------------------------ head management / drift regularisers
    def _get_base_weights(self):
        base_weight = self.classifier.weight.detach().clone().requires_grad_(False)
        if self.classifier.bias is not None:
            return base_weight, self.classifier.bias.detach().clone().requires_grad_(False)
        return base_weight, None

    def augment_base_classifier_(self, n, novel_weight=None, novel_bias=None):
        """Append n classifier rows; default rows come from a fresh nn.Linear(640, n) drawn on the CPU generator."""
        base_weight = self.classifier.weight.detach()
        base_bias = self.classifier.bias.detach() if self.classifier.bias is not None else None
        if novel_weight is None:
            novel_weight, drawn_bias = rng.linear_init(n, base_weight.size(1), base_bias is not None)
            if base_bias is not None and novel_bias is None:
                novel_bias = drawn_bias
        augmented = torch.cat([base_weight, novel_weight.to(base_weight.device)], 0)
        self.classifier.weight = nn.Parameter(augmented, requires_grad=True)
        if base_bias is not None:
            self.classifier.bias = nn.Parameter(torch.cat([base_bias, novel_bias.to(base_bias.device)]), requires_grad=True)

    def regloss(self, lmbd, base_weight, base_bias=None):
        reg = sra.scaled_dist(float(lmbd), self.classifier.weight[:base_weight.size(0), :], base_weight)
        if base_bias is not None:
            reg = reg + sra.scaled_sqdist(float(lmbd), self.classifier.bias[:base_weight.size(0)], base_bias)
        return reg

    def reglossnovel(self, lmbd, novel_weight, novel_bias=None):
        rng1, rng2 = self.num_classes, self.num_classes + novel_weight.size(0)
        reg = sra.scaled_dist(float(lmbd), self.classifier.weight[rng1:rng2, :], novel_weight)
        if novel_bias is not None:
            reg = reg + sra.scaled_sqdist(float(lmbd), self.classifier.bias[rng1:rng2], novel_bias)
        return reg


def resnet12(keep_prob=1.0, avg_pool=False, **kwargs):
    return ResNet(BasicBlock, [1, 1, 1, 1], keep_prob=keep_prob, avg_pool=avg_pool, **kwargs)


def resnet18(keep_prob=1.0, avg_pool=False, **kwargs):
    return ResNet(BasicBlock, [1, 1, 2, 2], keep_prob=keep_prob, avg_pool=avg_pool, **kwargs)
