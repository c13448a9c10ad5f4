"""Image transforms of the incremental evaluation (reference dataset/transform_cfg.py).

Only the deterministic tail of the test-time transform is on the hot path: ToTensor (uint8 HWC -> x / 255) and
Normalize(mean, std), with the constants of :8-10.  Here both are fused into the first kernel of the backbone
(`sr_pack_input_u8`): pass uint8 NHWC CUDA images (`MetaImageNet(..., raw=True)`) to `ResNet.features` /
`BackboneEngine.eval_features` instead of normalised fp32 NCHW tensors.

`transforms_test_options` / `transforms_options` / `transforms_list` keep `eval_incremental.py:18,50` working when this
package shadows the reference's `dataset`: option 'A' (miniImageNet) is the (support, query) pair of the reference -
random crop with 8-pixel padding + horizontal flip (+ colour jitter in the training variant) on the support copies,
plain normalisation on the queries - built from torchvision.  Without torchvision the entries are placeholders that
RAISE when they are applied (a silently dropped augmentation would turn the 5x tiled support set into five identical
copies).  The composed transforms carry `srb_kind` ('crop_flip' / 'crop_jitter_flip' / 'plain') so that the uint8 front
end (`MetaImageNet(raw=True)`) can apply the same crop / flip draws without PIL.  The CIFAR option 'D' belongs to datasets
outside this path.
"""
mean = [120.39586422 / 255.0, 115.59361427 / 255.0, 104.54012653 / 255.0]
std = [70.68188272 / 255.0, 68.27635443 / 255.0, 72.54505529 / 255.0]


def draw_crop_flip(n, size=84, padding=8, p=0.5):
    """Parameters of the reference's test-time SUPPORT augmentation for n images, drawn from torch's CPU generator in the
    order torchvision draws them per image - RandomCrop.get_params (two torch.randint) then RandomHorizontalFlip (one
    torch.rand) - so that `sr_pack_input_u8(crop_ij, flip, pad=padding)` reproduces `transforms_test_options['A'][0]` bit
    for bit while leaving the generator in the same state.  -> (int32 [n,2] crop corner in the padded image, uint8 [n])."""
    import torch
    ij = torch.empty((n, 2), dtype=torch.int32)
    flip = torch.empty(n, dtype=torch.uint8)
    span = 2 * padding + 1            # (size + 2 * padding) - size + 1 possible corners per axis
    for k in range(n):
        ij[k, 0] = torch.randint(0, span, size=(1,)).item()
        ij[k, 1] = torch.randint(0, span, size=(1,)).item()
        flip[k] = 1 if torch.rand(1) < p else 0
    return ij, flip


class _NeedsTorchvision(object):
    """Placeholder for a transform that cannot be built: fails loudly when used instead of dropping the augmentation."""

    def __init__(self, kind):
        self.srb_kind = kind

    def __call__(self, img):
        raise RuntimeError("srb200: the '%s' image transform needs torchvision + PIL (reference dataset/transform_cfg.py:"
                           "32-45); install them or use MetaImageNet(raw=True), whose uint8 front end needs neither" % self.srb_kind)


def apply_crop_flip_u8(batch_u8_nhwc, crop_ij, flip, padding=8):
    """RandomCrop(size, padding) + RandomHorizontalFlip on uint8 NHWC images with the given draws (see draw_crop_flip):
    the same pixels PIL produces (constant zero fill), without PIL."""
    import numpy as np
    x = np.asarray(batch_u8_nhwc)
    n, h, w, _ = x.shape
    padded = np.zeros((n, h + 2 * padding, w + 2 * padding, x.shape[3]), dtype=x.dtype)
    padded[:, padding:padding + h, padding:padding + w] = x
    out = np.empty_like(x)
    for k in range(n):
        i, j = int(crop_ij[k, 0]), int(crop_ij[k, 1])
        img = padded[k, i:i + h, j:j + w]
        out[k] = img[:, ::-1] if int(flip[k]) else img
    return out


def _option_a(jitter):
    kind = 'crop_jitter_flip' if jitter else 'crop_flip'
    try:
        import numpy as np
        import torchvision.transforms as T
        from PIL import Image
    except ImportError:
        import warnings
        warnings.warn("srb200: torchvision / PIL are missing - transforms_options['A'] will raise when applied to an image")
        return [_NeedsTorchvision(kind), _NeedsTorchvision('plain')]
    norm = T.Normalize(mean=mean, std=std)
    aug = [T.RandomCrop(84, padding=8)] + ([T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.4)] if jitter else []) + \
          [T.RandomHorizontalFlip()]
    support = T.Compose([Image.fromarray] + aug + [np.array, T.ToTensor(), norm])
    query = T.Compose([Image.fromarray, T.ToTensor(), norm])
    support.srb_kind, query.srb_kind = kind, 'plain'
    return [support, query]


transforms_list = ['A']
transforms_options = {'A': _option_a(jitter=True)}
transforms_test_options = {'A': _option_a(jitter=False)}
