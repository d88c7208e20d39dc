"""MIL encoders with the reference's constructor/forward surface and state_dict layout (model/backbone.py),
executed by the fused CUDA path.  ABMIL and DeepAttMISL are on the AdvMIL hot path; PatchGCN and the ESAT transformer
(`patch`) are out of scope for this build and raise NotImplementedError.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import get_precision, ops
from ..utils.func import next_dropout_seed
from .backbone_utils import Attn_Net_Gated


def load_backbone_param(mode, dims):
    if mode == "cluster":
        return [dims[:3]], {"num_clusters": 8, "dropout": 0.25}
    if mode in ("patch", "graph"):
        raise NotImplementedError(f"backbone mode '{mode}' is outside the B200 hot path (abmil / cluster are built)")
    return [dims[:3]], {"dropout": 0.25}


def Model_Zoo(mode):
    if mode == "cluster":
        return DeepAttMISL
    if mode in ("patch", "graph"):
        raise NotImplementedError(f"backbone mode '{mode}' is outside the B200 hot path (abmil / cluster are built)")
    return ABMIL


def load_backbone(mode, dims):
    """Same contract as the reference's load_backbone (model/backbone.py:47-51)."""
    net = Model_Zoo(mode)
    args, kws = load_backbone_param(mode, dims)
    return net(*args, **kws)


class _GatedMILBase(nn.Module):
    """Shared plumbing: collects the 14 generator tensors (GEN_TENSORS order) and runs the fused path."""

    kind = "abmil"

    def gen_params(self, head: Optional[nn.ModuleList] = None):
        an = self.attention_net
        gate = an[3]
        rho = getattr(self, "rho", None)
        p = [an[0].weight, an[0].bias, gate.attention_a[0].weight, gate.attention_a[0].bias, gate.attention_b[0].weight,
             gate.attention_b[0].bias, gate.attention_c.weight, gate.attention_c.bias,
             None if rho is None else rho[0].weight, None if rho is None else rho[0].bias]
        if head is None:
            p += [None, None, None, None]
        else:
            p += [head[0][0].weight, head[0][0].bias, head[1][0].weight, head[1][0].bias]
        return p


class ABMIL(_GatedMILBase):
    """Attention MIL encoder (reference model/backbone.py:54-86): attention_net = [Linear, ReLU, Dropout,
    Attn_Net_Gated], rho = [Linear, ReLU, Dropout]."""

    def __init__(self, dims: List, dropout: float = 0.25):
        super().__init__()
        assert len(dims) == 3
        dim_in, dim_hid, dim_out = dims
        self.dims = (dim_in, dim_hid, dim_out)
        self.p = dropout
        self.attention_net = nn.Sequential(nn.Linear(dim_in, dim_hid), nn.ReLU(), nn.Dropout(dropout),
                                           Attn_Net_Gated(L=dim_hid, D=dim_hid, dropout=dropout, n_classes=1))
        self.rho = nn.Sequential(nn.Linear(dim_hid, dim_out), nn.ReLU(), nn.Dropout(dropout))

    def config(self, hid=0, noise=(0, 0), out_scale=0, p_head=0.0) -> ops.GenConfig:
        C, h, o = self.dims
        return ops.GenConfig(C=C, h=h, o=o, hid=hid, noise0=noise[0], noise1=noise[1], out_scale=out_scale,
                             p_backbone=self.p, p_head=p_head)

    def forward(self, x_path, *args):
        """x_path [1,N,C] -> H [1,dim_out] (backbone-only mode of the fused generator kernels)."""
        bags = ops.PackedBags.from_single(x_path)
        H = ops.GeneratorFn.apply(self.config(), bags, None, None, None, self.training,
                                  next_dropout_seed() if self.training else 0, getattr(self, "_inject_masks", None),
                                  ops.PRECISIONS[get_precision()], *self.gen_params())
        return H


class DeepAttMISL(_GatedMILBase):
    """Cluster-based encoder (reference model/backbone.py:89-123): phis (1x1 conv == linear) + ReLU per instance,
    mean per cluster id (zeros for an empty cluster), then Linear+ReLU+Dropout, gated attention over the clusters."""

    kind = "cluster"

    def __init__(self, dims: List, num_clusters=8, dropout=0.25):
        super().__init__()
        assert len(dims) == 3
        dim_in, dim_hid, dim_out = dims
        assert dim_hid == dim_out
        self.dims = (dim_in, dim_hid, dim_out)
        self.dim_hid = dim_hid
        self.num_clusters = num_clusters
        self.p = dropout
        self.phis = nn.Sequential(nn.Conv2d(dim_in, dim_hid, 1), nn.ReLU())
        self.pool1d = nn.AdaptiveAvgPool1d(1)
        self.attention_net = nn.Sequential(nn.Linear(dim_hid, dim_hid), nn.ReLU(), nn.Dropout(dropout),
                                           Attn_Net_Gated(L=dim_hid, D=dim_hid, dropout=dropout, n_classes=1))

    def config(self, hid=0, noise=(0, 0), out_scale=0, p_head=0.0) -> ops.GenConfig:
        _, h, _ = self.dims
        # the attention stage sees "bags" of num_clusters rows of width h; no rho layer (H = w @ g)
        return ops.GenConfig(C=h, h=h, o=h, hid=hid, noise0=noise[0], noise1=noise[1], out_scale=out_scale,
                             p_backbone=self.p, p_head=p_head, has_rho=False)

    def cluster_rows(self, x_path: torch.Tensor, cluster_id: torch.Tensor, lengths=None) -> torch.Tensor:
        """[rows,C] (+ ids) -> per-(bag, cluster) mean of relu(phis(x)) [bags*num_clusters, h], differentiable."""
        x = x_path[0] if x_path.dim() == 3 else x_path
        cid = cluster_id.reshape(-1)  # accepts the handler's [1,N] as well as [N] (reference quirk A.4#7)
        assert cid.shape[0] == x.shape[0], "one cluster id per instance (dataset/PatchWSI.py:93)"
        lengths = [x.shape[0]] if lengths is None else lengths
        W = self.phis[0].weight
        return ClusterPoolFn.apply(x, cid.to(torch.int32), lengths, self.num_clusters,
                                   ops.PRECISIONS[get_precision()], W.reshape(W.shape[0], -1), self.phis[0].bias)

    def forward(self, x_path, cluster_id, *args):
        hc = self.cluster_rows(x_path, cluster_id)
        bags = ops.PackedBags(hc, [self.num_clusters])
        H = ops.GeneratorFn.apply(self.config(), bags, hc, None, None, self.training,
                                  next_dropout_seed() if self.training else 0, getattr(self, "_inject_masks", None),
                                  ops.FP32, *self.gen_params())   # num_clusters rows: always the exact fp32 engine
        return H


class ClusterPoolFn(torch.autograd.Function):
    """K9: hx = relu(x Wphi^T + b) on every row (projection GEMM), then segment-mean by cluster id."""

    @staticmethod
    def forward(ctx, x, cid, lengths, ncl, precision, W, b):
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        bags = ops.PackedBags(x.detach(), lengths).for_precision(precision)
        hx = ops.linear_forward(bags.x, W.detach(), b.detach(), act=1, precision=precision)
        elem = ops.ELEM_BF16 if hx.dtype == torch.bfloat16 else ops.ELEM_F32
        h = W.shape[0]
        out = torch.empty(bags.bags * ncl, h, dtype=torch.float32, device=x.device)
        counts = torch.empty(bags.bags * ncl, dtype=torch.int32, device=x.device)
        cid = cid.contiguous()
        wsb = lib.advmil_segment_mean_workspace_bytes(bags.rows, bags.bags, h, ncl)
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=x.device)
        _lib.check(lib.advmil_segment_mean_by_id_fwd(hx.data_ptr(), elem, cid.data_ptr(), bags.offsets.data_ptr(), bags.offsets_host,
                                                     bags.rows, bags.bags, h, ncl, out.data_ptr(), counts.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                   "advmil_segment_mean_by_id_fwd")
        ctx.saved = (bags, hx, cid, counts, W, ncl, precision)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from .. import _lib
        lib = _lib.load()
        bags, hx, cid, counts, W, ncl, precision = ctx.saved
        h = W.shape[0]
        d_hx = torch.empty_like(hx)
        d_out = d_out.contiguous().float()
        elem = ops.ELEM_BF16 if hx.dtype == torch.bfloat16 else ops.ELEM_F32
        _lib.check(lib.advmil_segment_mean_by_id_bwd(d_out.data_ptr(), hx.data_ptr(), elem, cid.data_ptr(), bags.offsets.data_ptr(),
                                                     counts.data_ptr(), bags.rows, bags.bags, h, ncl, 1, d_hx.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "advmil_segment_mean_by_id_bwd")
        _, dW, db = ops.linear_backward(d_hx, bags.x, W.detach(), need_dx=False, precision=precision)
        return None, None, None, None, None, dW.reshape(W.shape), db
