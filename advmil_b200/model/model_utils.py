"""Builders and containers mirroring the reference's model/model_utils.py surface (names, shapes, init)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .backbone_utils import GAPool, make_embedding_layer


@torch.no_grad()
def init_weights(m):
    """Xavier-uniform weights and zero bias on every nn.Linear (reference model/model_utils.py:12-17); the handler
    applies it to the generator only (model/model_handler.py:81)."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            m.bias.data.zero_()


def get_hop_dims(d, hops):
    res, cur = [], d
    for _ in range(hops):
        cur = cur // 2
        if cur <= 1:
            break
        res.append(cur)
    return res


def make_mlp_layer(dim_in, dim_out, layer_norm=True, dropout=0.25):
    layers = [nn.Linear(dim_in, dim_out), nn.ReLU(inplace=True), nn.Dropout(dropout)]
    if layer_norm:
        layers.insert(1, nn.LayerNorm(dim_out))
    return nn.Sequential(*layers)


def make_efficient_mlp_layer(dim, layer_norm=True, dropout=0.25):
    if layer_norm:
        # the reference's layer_norm=True branch references an undefined name (model/model_utils.py:165); only False is reachable
        raise NotImplementedError("make_efficient_mlp_layer(layer_norm=True) is unreachable in the reference")
    return nn.Sequential(nn.Linear(dim, dim // 2), nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(dim // 2, dim))


def make_noise_mlp_layer(in_dim: int, out_dim: int, noise, hops: int = 1, norm: bool = False, dropout: float = 0.25):
    """Noise-concat MLP stack (reference model/model_utils.py:116-133): hidden dims halve per hop; a layer whose
    noise flag is 1 takes [H, N] of twice the width; the last layer is a plain Linear."""
    hid = get_hop_dims(in_dim, hops)
    ins, outs = [in_dim] + hid, hid + [out_dim]
    mlps = nn.ModuleList()
    for i in range(len(hid) + 1):
        din = ins[i] * (2 if noise[i] == 1 else 1)
        if i == len(hid):
            mlps.append(nn.Sequential(nn.Linear(din, outs[i])))
        else:
            mlps.append(make_mlp_layer(din, outs[i], norm, dropout))
    return mlps


def make_embedding_y_layer(args):
    """[B,1] -> [B,C'] time embedding: a (Linear, ReLU, Dropout) block per hidden dim (reference :178-186)."""
    in_dim, layers = args.in_dim, []
    for hd in args.hid_dims:
        layers.append(make_mlp_layer(in_dim, hd, args.norm, args.dropout))
        in_dim = hd
    return nn.Sequential(*layers)


class EmbedXLayer(nn.Module):
    """[B,N,C] -> region embeddings -> fc1 -> GAPool -> fc2 (reference model/model_utils.py:188-210).
    Parameter container; PrjDiscriminator runs it through the fused RLIP kernels."""

    def __init__(self, args):
        super().__init__()
        out_dim = args.out_dim
        args.scale = 4
        args.dw_conv = False
        self.embedding = make_embedding_layer(args.backbone, args)
        self.fc1 = make_efficient_mlp_layer(out_dim, False, args.dropout)
        self.pool = GAPool(out_dim, out_dim, args.dropout)
        self.fc2 = make_efficient_mlp_layer(out_dim, False, args.dropout)
        self.p = args.dropout

    def disc_params(self):
        e = self.embedding
        return [e.conv.weight, e.conv.bias, e.norm.weight, e.norm.bias,
                self.fc1[0].weight, self.fc1[0].bias, self.fc1[3].weight, self.fc1[3].bias,
                self.pool.fc1[0].weight, self.pool.fc1[0].bias, self.pool.score[0].weight, self.pool.score[0].bias,
                self.pool.fc2.weight, self.pool.fc2.bias,
                self.fc2[0].weight, self.fc2[0].bias, self.fc2[3].weight, self.fc2[3].bias]
