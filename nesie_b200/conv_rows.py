"""1x1 Conv1d (+ BatchNorm1d + ReLU) stacks of the head evaluated as row GEMMs on this repo's kernels."""
import torch
from torch import nn as nn
from torch.nn import functional as F

from . import bn_rows
from . import mlp_rows
from .pool_rows import add_bias_rows
from .pointnet_modules import ConvModule, _fused_bn, _rows_linear


def conv1d_rows(seq, x):
    """Apply a stack of 1x1 Conv1d (+BN1d+ReLU) modules to (B, C, n) as row GEMMs on (B*n, C):
    the same parameters and arithmetic as the Conv1d modules (cuDNN's fp32 1x1 wgrad engines are
    slow), used on CUDA tensors; plain module calls otherwise (CPU oracle twin)."""
    mods = list(seq) if isinstance(seq, nn.Sequential) else [seq]
    if not x.is_cuda:
        for m in mods:
            x = m(x)
        return x
    B, C, n = x.shape
    r = x.transpose(1, 2).reshape(B * n, C)
    if _fused_bn() and all(isinstance(m, ConvModule) and m.conv.bias is None for m in mods):
        pairs = [(m.conv.weight.flatten(1), m.bn) for m in mods]
        if mlp_rows.supported(r, pairs):  # BatchNorm fused into the GEMMs (mlp_rows.py)
            return mlp_rows.mlp_rows(r, pairs).view(B, n, -1).transpose(1, 2)
    for m in mods:
        conv = m.conv if isinstance(m, ConvModule) else m
        r = _rows_linear(r, conv.weight.flatten(1))
        if conv.bias is not None:
            r = add_bias_rows(r, conv.bias)
        if isinstance(m, ConvModule):
            bn = m.bn
            if _fused_bn() and bn_rows.supported(r, bn):
                r = bn_rows.bn_relu_rows(r, bn)
                continue
            if bn.training and bn.track_running_stats:
                bn_rows.count_batch(bn)
            r = F.batch_norm(r, bn.running_mean, bn.running_var, bn.weight, bn.bias,
                             bn.training or not bn.track_running_stats, bn.momentum, bn.eps)
            r = F.relu(r, inplace=True)
    return r.view(B, n, -1).transpose(1, 2)
