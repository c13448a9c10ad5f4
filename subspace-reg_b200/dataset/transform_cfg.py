"""Image transforms of the incremental evaluation (reference dataset/transform_cfg.py).

Only the deterministic tail of the test-time transform is on the hot path: ToTensor (uint8 HWC -> x / 255) and
Normalize(mean, std), with the constants of :8-10.  Here both are fused into the first kernel of the backbone
(`sr_pack_input_u8`): pass uint8 NHWC CUDA images (`MetaImageNet(..., raw=True)`) to `ResNet.features` /
`BackboneEngine.eval_features` instead of normalised fp32 NCHW tensors.

`transforms_test_options` / `transforms_options` / `transforms_list` keep `eval_incremental.py:18,50` working when this
package shadows the reference's `dataset`: option 'A' (miniImageNet) is the (support, query) pair of the reference -
random crop with 8-pixel padding + horizontal flip (+ colour jitter in the training variant) on the support copies,
plain normalisation on the queries - built from torchvision when it is installed; without torchvision the pair is
(None, None), which `dataset.mini_imagenet` reads as "normalise only" on both branches.  The CIFAR option 'D' belongs to
datasets outside this path.
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


def _option_a(jitter):
    try:
        import numpy as np
        import torchvision.transforms as T
        from PIL import Image
    except ImportError:
        return [None, None]
    norm = T.Normalize(mean=mean, std=std)
    aug = [T.RandomCrop(84, padding=8)] + ([T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.4)] if jitter else []) + \
          [T.RandomHorizontalFlip()]
    support = T.Compose([Image.fromarray] + aug + [np.array, T.ToTensor(), norm])
    query = T.Compose([Image.fromarray, T.ToTensor(), norm])
    return [support, query]


transforms_list = ['A']
transforms_options = {'A': _option_a(jitter=True)}
transforms_test_options = {'A': _option_a(jitter=False)}
