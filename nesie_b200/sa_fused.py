"""Host side of the fused tcgen05 set-abstraction forward (eval mode, folded BatchNorm, bf16).

Covers the chain QueryAndGroup body -> 3 x ConvModule(1x1 conv, BN, ReLU) -> max-pool of the
reference's PointSAModule (ops/pointnet_modules/point_sa_module.py:191-211) in ONE kernel,
nesie_sa_fused_forward.  This module folds the BatchNorm running statistics into per-channel
scale/shift, packs the three weight matrices into byte images of the UMMA shared-memory operand
layout (K-major, 128-byte swizzle) and converts the (B, C, N) fp32 features into the bf16
point-major table the kernel gathers from.
"""
import torch

from . import _lib


def supported(nsample, c_in, c1, c2, c3):
    return bool(_lib.lib().nesie_sa_fused_supported(nsample, c_in, c1, c2, c3))


def pack_weight_image(w, kpad):
    """(C_out, K) fp32 -> uint8 image [ceil(kpad/64) slabs][C_out rows][128 B]: bf16, K-major,
    16-byte chunk c of row r stored at chunk position c ^ (r & 7) (SWIZZLE_128B)."""
    cout, k = w.shape
    nslab = (kpad + 63) // 64
    wp = torch.zeros((cout, nslab * 64), dtype=torch.float32, device=w.device)
    wp[:, :k] = w
    wb = wp.to(torch.bfloat16).view(cout, nslab, 8, 8)            # row, slab, chunk, 8 elems
    rows = torch.arange(cout, device=w.device).view(cout, 1, 1)
    chunk = torch.arange(8, device=w.device).view(1, 1, 8)
    dst_chunk = (chunk ^ (rows & 7)).expand(cout, nslab, 8)         # where each chunk goes
    img = torch.empty((nslab, cout, 8, 8), dtype=torch.bfloat16, device=w.device)
    src = wb.permute(1, 0, 2, 3).contiguous()                       # slab, row, chunk, elem
    img.scatter_(2, dst_chunk.permute(1, 0, 2).unsqueeze(-1).expand(nslab, cout, 8, 8), src)
    return img.contiguous().view(torch.uint8).reshape(-1)


def fold_mlp(mlp, c_in):
    """ConvModule x3 (conv 1x1 no bias + BN, eval statistics) -> packed kernel arguments."""
    layers = list(mlp.children())
    assert len(layers) == 3, "the fused kernel covers three-layer shared MLPs"
    c8 = 8 if c_in == 0 else (c_in + 7) // 8 * 8
    k0pad = (c8 + 3 + 15) // 16 * 16
    imgs, ss, dims = [], [], []
    for li, layer in enumerate(layers):
        w = layer.conv.weight.detach().float().flatten(1)           # (C_out, C_in_total)
        bn = layer.bn
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
        shift = bn.bias.detach().float() - bn.running_mean.float() * scale
        if layer.conv.bias is not None:
            shift = shift + layer.conv.bias.detach().float() * scale
        if li == 0:  # reference channel order is [xyz(3), feats]; the kernel's is [feats | xyz]
            w1 = torch.zeros((w.shape[0], k0pad), dtype=torch.float32, device=w.device)
            w1[:, :c_in] = w[:, 3:3 + c_in]
            w1[:, c8:c8 + 3] = w[:, :3]
            w, kpad = w1, k0pad
        else:
            kpad = w.shape[1]
        imgs.append(pack_weight_image(w * scale[:, None], kpad))  # BN scale folded into W
        ss.append(shift)
        dims.append(w.shape[0])
    return dict(w1=imgs[0], w2=imgs[1], w3=imgs[2], scale_shift=torch.cat(ss).contiguous(),
                c_in=c_in, c1=dims[0], c2=dims[1], c3=dims[2])


def pack_features(features, B, N):
    """(B, C, N) fp32 or None -> (B, N, c8) bf16 point-major table."""
    dev = features.device if features is not None else None
    c = 0 if features is None else features.shape[1]
    c8 = 8 if c == 0 else (c + 7) // 8 * 8
    if features is None:
        raise ValueError("pack_features needs a device; pass zeros for xyz-only SA modules")
    table = torch.empty((B, N, c8), dtype=torch.bfloat16, device=dev)
    features_c = features.contiguous()   # alive across the launch
    with torch.cuda.device(dev):
        _lib.call("nesie_pack_features_bf16", B, c, N, _lib.ptr(features_c),
                  _lib.ptr(table), _lib.stream())
    return table


def sa_fused_forward(points_xyz, center_xyz, features, idx, radius, packed):
    """points_xyz (B,N,3), center_xyz (B,M,3), features (B,C,N) fp32, idx (B,M,K) int32,
    radius > 0 to normalise (0 = off), packed = fold_mlp(...)  ->  (B, C3, M) fp32."""
    _lib.need_cuda(points_xyz, center_xyz, features, idx)
    B, N, _ = points_xyz.shape
    M, K = idx.shape[1], idx.shape[2]
    table = pack_features(features, B, N)
    out = torch.empty((B, packed['c3'], M), dtype=torch.float32, device=points_xyz.device)
    xyz_c, center_c, idx = points_xyz.contiguous(), center_xyz.contiguous(), idx.contiguous()
    with torch.cuda.device(points_xyz.device):
        _lib.call("nesie_sa_fused_forward", B, N, M, K, packed['c_in'], packed['c1'], packed['c2'],
                  packed['c3'], _lib.ptr(xyz_c), _lib.ptr(center_c),
                  _lib.ptr(table), _lib.ptr(idx), float(radius), _lib.ptr(packed['w1']),
                  _lib.ptr(packed['w2']), _lib.ptr(packed['w3']), _lib.ptr(packed['scale_shift']),
                  _lib.ptr(out), _lib.stream())
    return out
