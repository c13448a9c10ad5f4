"""Normalisation constants of the reference's image pipeline (dataset/transform_cfg.py:8-10).

Only the deterministic tail of its test-time transform is on the incremental-session path: ToTensor (uint8 HWC -> x / 255)
and Normalize(mean, std).  Here both are fused into the first kernel of the backbone (`sr_pack_input_u8`): pass uint8
NHWC CUDA images to `ResNet.features` / `BackboneEngine.eval_features` instead of normalised fp32 NCHW tensors.  The
PIL / torchvision augmentations of the support set (RandomCrop, ColorJitter, flip) stay outside.
"""
mean = [120.39586422 / 255.0, 115.59361427 / 255.0, 104.54012653 / 255.0]
std = [70.68188272 / 255.0, 68.27635443 / 255.0, 72.54505529 / 255.0]
