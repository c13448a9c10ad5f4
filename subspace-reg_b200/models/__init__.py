"""Drop-in for the reference's ``models`` package (models/__init__.py:1-11): same names, B200-native internals."""
from .resnet_language import resnet12, resnet18

model_pool = [
    'resnet12',
    'resnet18',
]

model_dict = {
    'resnet12': resnet12,
    'resnet18': resnet18,
}
