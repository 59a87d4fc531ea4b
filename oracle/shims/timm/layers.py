"""Shim for the un-vendored `timm` dependency of the reference (pe.py:14).
drop_path is 0.0 in every PE config, so DropPath is the identity."""
import torch.nn as nn


class DropPath(nn.Identity):
    def __init__(self, *a, **k):
        super().__init__()
