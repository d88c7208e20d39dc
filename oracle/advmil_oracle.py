"""CPU oracle for the AdvMIL G+D hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the reference algorithm (liupei101/AdvMIL) in plain
PyTorch-on-CPU functional form.  It is not the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.
The product path (advmil_b200.*) never imports anything from `oracle/`.

Parity pin: the reference ships no tests and no golden vectors ("parity unpinned" by the
reference's own tests).  The pin used here is the *live reference itself*:
`oracle/make_golden.py` imports the reference modules from /root/reference in the build container,
runs them on seeded inputs and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
this restatement against those fixtures (max abs err ~1e-7 in fp32, exact in index work).

Every function cites the reference file:line it follows (paths relative to the reference root).
State dicts use the reference's parameter names so checkpoints interchange.

All functions are differentiable torch code; gradients for parity come from autograd on fp32
(or fp64 when `dtype=torch.float64` inputs are given).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _drop(v: Tensor, mask: Optional[Tensor], p: float) -> Tensor:
    """nn.Dropout(p) in train mode with an explicit keep-mask (1 = keep). mask None = eval mode."""
    if mask is None or p <= 0.0:
        return v
    return v * mask.to(v.dtype) / (1.0 - p)


def _lin(v: Tensor, sd: SD, name: str) -> Tensor:
    return F.linear(v, sd[name + ".weight"], sd.get(name + ".bias"))


# ---- optional emulation of the bf16 STORAGE mode of the CUDA path (tests only) ----------------------------------
# The reference is fp32 end to end.  The CUDA path's bf16 mode keeps x, the N-row GEMM operands (weights as read by the
# tensor cores) and the N-row activations / activation gradients it writes to HBM (h, y_pre, dAB, dh, dy) in bfloat16,
# with fp32 accumulation, statistics, parameters and gradients.  Under `with bf16_storage():` the N-row stages below
# round at the same points (straight-through), so that a kernel bug cannot hide behind the ReLU-mask flips that ANY
# bf16 evaluation shows against the fp32 reference.  Off by default: the oracle proper is the fp32 restatement.
_EMU_BF16 = False


class bf16_storage:
    def __enter__(self):
        global _EMU_BF16
        self.prev, _EMU_BF16 = _EMU_BF16, True

    def __exit__(self, *a):
        global _EMU_BF16
        _EMU_BF16 = self.prev


def _round_bf16(v: Tensor) -> Tensor:
    return v.to(torch.bfloat16).to(v.dtype)


class _RoundSTE(torch.autograd.Function):     # forward: round to bf16; backward: identity
    @staticmethod
    def forward(ctx, v):
        return _round_bf16(v)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):    # forward: identity; backward: gradient rounded to bf16 (dAB / dh / dy)
    @staticmethod
    def forward(ctx, v):
        return v.view_as(v)

    @staticmethod
    def backward(ctx, g):
        return _round_bf16(g)


def _q(v: Tensor) -> Tensor:
    return _RoundSTE.apply(v) if _EMU_BF16 else v


def _qg(v: Tensor) -> Tensor:
    return _RoundGrad.apply(v) if _EMU_BF16 else v


def _lin_rows(v: Tensor, W: Tensor, b: Optional[Tensor]) -> Tensor:
    """An N-row contraction: in emulation mode both tensor-core operands are bf16 and the gradient flowing back into the
    pre-activation is stored as bf16."""
    return _qg(F.linear(_q(v), _q(W), b))


# ----------------------------------------------------------------------------------------------
# G: ABMIL backbone + gated attention + pooling  (model/backbone.py:54-86, backbone_utils.py:11-29)
# ----------------------------------------------------------------------------------------------
def gated_attention_scores(h: Tensor, sd: SD, prefix: str, masks=None, p: float = 0.25) -> Tensor:
    """Attn_Net_Gated.forward (model/backbone_utils.py:24-29): s = (tanh(hWa+ba) * sig(hWb+bb)) wc + bc."""
    masks = masks or {}
    rows = _lin_rows if (_EMU_BF16 and prefix.startswith("backbone")) else F.linear   # GAPool of D is region-level: fp32
    a = _drop(torch.tanh(rows(h, sd[prefix + ".attention_a.0.weight"], sd[prefix + ".attention_a.0.bias"])), masks.get("a"), p)
    b = _drop(torch.sigmoid(rows(h, sd[prefix + ".attention_b.0.weight"], sd[prefix + ".attention_b.0.bias"])), masks.get("b"), p)
    return _lin(a * b, sd, prefix + ".attention_c")  # [N, 1]


def abmil_forward(sd: SD, x: Tensor, masks=None, p: float = 0.25, prefix: str = "backbone") -> Dict[str, Tensor]:
    """ABMIL.forward (model/backbone.py:79-86). x: [N, C] (the reference squeezes [1,N,C]).

    masks: optional dict of keep-masks {'h':[N,h], 'a':[N,h], 'b':[N,h], 'rho':[1,h]} = train mode.
    Returns h (post-dropout, the tensor that is pooled — quirk A.4#4), s, w, z, H.
    """
    masks = masks or {}
    h = _drop(torch.relu(_lin_rows(x, sd[f"{prefix}.attention_net.0.weight"], sd[f"{prefix}.attention_net.0.bias"])),
              masks.get("h"), p)                                                           # backbone.py:67-70
    h = _q(h)
    s = gated_attention_scores(h, sd, f"{prefix}.attention_net.3", masks, p)              # backbone.py:71
    w = torch.softmax(s.transpose(1, 0), dim=1)                                            # backbone.py:82-83
    z = w @ h                                                                              # backbone.py:84
    H = _drop(torch.relu(_lin(z, sd, f"{prefix}.rho.0")), masks.get("rho"), p)             # backbone.py:73-77,85
    return {"h": h, "s": s.squeeze(-1), "w": w.squeeze(0), "z": z, "H": H}


def deepattmisl_forward(sd: SD, x: Tensor, cluster_id: Tensor, num_clusters: int = 8, masks=None,
                        p: float = 0.25, prefix: str = "backbone") -> Dict[str, Tensor]:
    """DeepAttMISL.forward (model/backbone.py:105-123).

    Per cluster c: mean over {n: cid[n]==c} of relu(x_n Wphi^T + bphi) (1x1 conv == linear,
    backbone.py:98,112-116), zeros for an empty cluster (:114-115); then Linear+ReLU+Dropout,
    gated attention over the clusters, softmax, weighted sum (:117-122). No rho layer.
    """
    masks = masks or {}
    x = x.reshape(-1, x.shape[-1])
    cid = cluster_id.reshape(-1)
    Wphi = sd[f"{prefix}.phis.0.weight"].reshape(sd[f"{prefix}.phis.0.weight"].shape[0], -1)
    hc = []
    for c in range(num_clusters):
        sel = cid == c
        if int(sel.sum()) == 0:
            hc.append(torch.zeros(Wphi.shape[0], dtype=x.dtype))
        else:
            hc.append(torch.relu(F.linear(x[sel], Wphi, sd[f"{prefix}.phis.0.bias"])).mean(dim=0))
    hc = torch.stack(hc, dim=0)                                                            # [8, h]
    g = _drop(torch.relu(_lin(hc, sd, f"{prefix}.attention_net.0")), masks.get("h"), p)
    s = gated_attention_scores(g, sd, f"{prefix}.attention_net.3", masks, p)
    w = torch.softmax(s.transpose(1, 0), dim=1)
    H = w @ g
    return {"hc": hc, "g": g, "s": s.squeeze(-1), "w": w.squeeze(0), "H": H}


def posemb_sincos_2d(y: Tensor, x: Tensor, dim: int, temperature: float = 10000.0) -> Tensor:
    """model/backbone_utils.py:79-88: [x.sin | x.cos | y.sin | y.cos] with dim/4 frequencies each."""
    assert dim % 4 == 0
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    yy = y.flatten()[:, None] * omega[None, :]
    xx = x.flatten()[:, None] * omega[None, :]
    return torch.cat((xx.sin(), xx.cos(), yy.sin(), yy.cos()), dim=1)


def compute_pe(coord: Tensor, ndim: int = 384, step: int = 1) -> Tensor:
    """model/backbone_utils.py:90-99 (+ to_relative_coord): coord [R,2] of the level-2 regions (after discretisation) ->
    PE [R, ndim] of the coordinates relative to their minimum."""
    ref_xy, _ = torch.min(coord, dim=-2)
    ncoord = coord - ref_xy
    y = torch.div(ncoord[:, 1], step, rounding_mode="floor")
    x = torch.div(ncoord[:, 0], step, rounding_mode="floor")
    return posemb_sincos_2d(y, x, ndim).to(torch.float32)


def encoder_layer(sd: SD, v: Tensor, prefix: str, nhead: int = 8, masks=None, p: float = 0.25, eps: float = 1e-5) -> Dict[str, Tensor]:
    """nn.TransformerEncoderLayer(d, nhead, dim_feedforward=d, dropout=p, activation='relu', batch_first=True) as built by
    make_transformer_layer (model/backbone_utils.py:112-127), post-norm: x1 = norm1(x + drop1(SA(x)));
    x2 = norm2(x1 + drop2(linear2(drop(relu(linear1(x1)))))).  SA = nn.MultiheadAttention: packed in_proj, scaled dot
    product (1/sqrt(d/nhead)), softmax over keys, dropout on the probabilities, out_proj.  v: [R, d] (one bag).
    masks: 'attn' [nhead,R,R], 'sa' [R,d], 'ff1' [R,d], 'ff2' [R,d]."""
    masks = masks or {}
    R, d = v.shape
    hd = d // nhead
    qkv = F.linear(v, sd[prefix + ".self_attn.in_proj_weight"], sd[prefix + ".self_attn.in_proj_bias"])
    q, k, val = [t.reshape(R, nhead, hd).transpose(0, 1) for t in qkv.split(d, dim=1)]       # [nhead, R, hd]
    scores = (q @ k.transpose(1, 2)) / math.sqrt(hd)
    P = _drop(torch.softmax(scores, dim=-1), masks.get("attn"), p)
    ctx = (P @ val).transpose(0, 1).reshape(R, d)
    sa = _drop(_lin(ctx, sd, prefix + ".self_attn.out_proj"), masks.get("sa"), p)
    x1 = F.layer_norm(v + sa, (d,), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"], eps)
    f = _drop(torch.relu(_lin(x1, sd, prefix + ".linear1")), masks.get("ff1"), p)
    f2 = _drop(_lin(f, sd, prefix + ".linear2"), masks.get("ff2"), p)
    x2 = F.layer_norm(x1 + f2, (d,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], eps)
    return {"ctx": ctx, "x1": x1, "x2": x2}


def esat_forward(sd: SD, x: Tensor, coord: Optional[Tensor] = None, masks=None, p: float = 0.25, nhead: int = 8,
                 prefix: str = "backbone") -> Dict[str, Tensor]:
    """DualTrans_HS.forward (model/backbone.py:189-196) with the defaults of load_backbone_param('patch')
    (:31-35): AVGPoolPatchEmbedding(1024 -> d, scale 4, ksize 1) -> (+ PE) -> 1 TransformerEncoderLayer(d, 8 heads,
    ffn d, dropout 0.25) -> GAPool(d, d).  x: [N, C], N % 16 == 0; coord: [R, 2] or None.
    masks (train mode): 'attn', 'sa', 'ff1', 'ff2' (encoder layer), 'ga', 'gs' (GAPool)."""
    emb = region_embed(sd, x, prefix + ".patch_embedding_layer")["emb"]                      # backbone.py:191
    if coord is not None:
        emb = emb + compute_pe(coord, emb.shape[1]).to(emb.dtype)                            # backbone.py:192-194
    enc = encoder_layer(sd, emb, prefix + ".patch_encoder_layer.layers.0", nhead, masks, p)  # backbone.py:195
    gp = gapool(sd, enc["x2"], prefix + ".pool", masks, p)                                   # backbone.py:196
    return {"emb": emb, "x1": enc["x1"], "x2": enc["x2"], "attn": gp["attn"], "H": gp["out"]}


def generator_head(sd: SD, H: Tensor, noises: Sequence[Optional[Tensor]], noise_cfg: Sequence[int],
                   masks=None, p: float = 0.6, out_scale: str = "sigmoid") -> Tensor:
    """Noise-concat MLP head: Generator.forward (model/GANSurv.py:32-49) over the layers built by
    make_noise_mlp_layer (model/model_utils.py:116-133): hidden layers Linear+ReLU+Dropout(p),
    last layer plain Linear; noise of the same width concatenated in front of layers with noise[i]==1.
    """
    masks = masks or {}
    n_layers = len(noise_cfg)
    for i in range(n_layers):
        data = torch.cat([H, noises[i]], dim=1) if noise_cfg[i] == 1 else H
        H = _lin(data, sd, f"MLPs.{i}.0")
        if i != n_layers - 1:
            H = _drop(torch.relu(H), masks.get(f"mlp{i}"), p)
    if out_scale == "sigmoid":
        return torch.sigmoid(H)
    if out_scale == "exp":
        return torch.exp(H)
    return H


def generator_forward(sd: SD, x: Tensor, noises, noise_cfg=(0, 1), masks=None, backbone: str = "abmil",
                      cluster_id: Optional[Tensor] = None, gen_dropout: float = 0.6,
                      out_scale: str = "sigmoid", coord: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Generator.forward (model/GANSurv.py:30-49). `noises[i]` is the tensor generate_noise would have
    produced for layer i (utils/func.py:154-164 == torch.rand / torch.randn on the CPU stream)."""
    if backbone == "cluster":
        bb = deepattmisl_forward(sd, x, cluster_id, masks=masks)
    elif backbone == "patch":
        bb = esat_forward(sd, x, coord, masks=masks)
    else:
        bb = abmil_forward(sd, x, masks=masks)
    pred = generator_head(sd, bb["H"], noises, noise_cfg, masks, gen_dropout, out_scale)
    bb["pred"] = pred
    return bb


# ----------------------------------------------------------------------------------------------
# D: region-level instance projection (RLIP)  (GANSurv.py:71-105, model_utils.py:157-210,
#    backbone_utils.py:31-77,129-168)
# ----------------------------------------------------------------------------------------------
def region_of_row(n: np.ndarray, scale: int = 4) -> np.ndarray:
    """Region id of level-1 row n: sequence2square (model/backbone_utils.py:62-69) views rows
    16r..16r+15 as region r; this is the inverse of tools/big_to_small_patching.py:40-46,70-72."""
    return np.asarray(n) // (scale * scale)


def region_embed(sd: SD, x: Tensor, prefix: str = "net_pair_one.embedding", eps: float = 1e-5) -> Dict[str, Tensor]:
    """AVGPoolPatchEmbedding.forward with ksize=1, stride=1, scale=4 (model/backbone_utils.py:158-168):
    emb[r] = mean_{k<16} relu(LayerNorm_128(x[16r+k] Wc^T + bc)).  x: [N, C], N % 16 == 0 (:65)."""
    N = x.shape[0]
    assert N % 16 == 0
    Wc = sd[f"{prefix}.conv.weight"]
    Wc = Wc.reshape(Wc.shape[0], -1)
    y = _q(_lin_rows(x, Wc, sd[f"{prefix}.conv.bias"]))
    e = torch.relu(F.layer_norm(y, (y.shape[-1],), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], eps))
    emb = e.reshape(N // 16, 16, -1).mean(dim=1)
    return {"y": y, "e": e, "emb": emb}


def _eff_mlp(v: Tensor, sd: SD, prefix: str, mask, p: float) -> Tensor:
    """make_efficient_mlp_layer(dim, False, p) (model/model_utils.py:157-166): Lin(d,d/2) ReLU Drop Lin(d/2,d)."""
    return _lin(_drop(torch.relu(_lin(v, sd, prefix + ".0")), mask, p), sd, prefix + ".3")


def gapool(sd: SD, v: Tensor, prefix: str, masks=None, p: float = 0.25) -> Dict[str, Tensor]:
    """GAPool.forward (model/backbone_utils.py:47-56) on [R, d] -> [1, d]; pools its *input* v."""
    masks = masks or {}
    ga = _drop(torch.tanh(_lin(v, sd, prefix + ".fc1.0")), masks.get("ga"), p)
    gs = _drop(torch.sigmoid(_lin(v, sd, prefix + ".score.0")), masks.get("gs"), p)
    rep = _lin(ga * gs, sd, prefix + ".fc2")              # [R,1]
    attn = torch.softmax(rep.transpose(1, 0), dim=1)      # [1,R]
    return {"rep": rep.squeeze(-1), "attn": attn.squeeze(0), "out": attn @ v}


def embedx_forward(sd: SD, x: Tensor, masks=None, p: float = 0.25, prefix: str = "net_pair_one") -> Dict[str, Tensor]:
    """EmbedXLayer.forward(x, return_instance=True) (model/model_utils.py:202-210).
    Returns hx [1,d] and fi [R,d] — NB the *post-fc1* features are the 'instance embeddings' (quirk A.4#2)."""
    masks = masks or {}
    re = region_embed(sd, x, prefix + ".embedding")
    fi = _eff_mlp(re["emb"], sd, prefix + ".fc1", masks.get("fc1"), p)
    gp = gapool(sd, fi, prefix + ".pool", masks, p)
    hx = _eff_mlp(gp["out"], sd, prefix + ".fc2", masks.get("fc2"), p)
    return {"emb": re["emb"], "fi": fi, "attn": gp["attn"], "bag": gp["out"], "hx": hx}


def time_embed(sd: SD, t: Tensor, n_layers: int = 2, prefix: str = "net_pair_two") -> Tensor:
    """make_embedding_y_layer with norm=False, dropout=0 (model/model_utils.py:178-186): (Linear+ReLU) x n."""
    v = t
    for i in range(n_layers):
        v = torch.relu(_lin(v, sd, f"{prefix}.{i}.0"))
    return v


def prjdisc_forward(sd: SD, x: Tensor, t: Tensor, masks=None, inner_product: str = "instance",
                    prj_path: str = "x", p: float = 0.25) -> Dict[str, Tensor]:
    """PrjDiscriminator.forward (model/GANSurv.py:89-105). x: [N,C]; t: [1,1]. Returns out [1,1]."""
    ht = time_embed(sd, t)
    ex = embedx_forward(sd, x, masks, p)
    if inner_product == "bag":
        out = (ht * ex["hx"]).sum(dim=-1, keepdim=True)                    # GANSurv.py:92-94
    else:
        out_ins = (ex["fi"].unsqueeze(0) * ht).sum(dim=-1)                 # GANSurv.py:97  [1,R]
        out = out_ins.mean(dim=-1, keepdim=True)                           # GANSurv.py:98
    if prj_path == "x":
        out = out + _lin(ex["hx"], sd, "prj_layer")                        # GANSurv.py:102-104
    elif prj_path == "y":
        out = out + _lin(ht, sd, "prj_layer")
    ex["ht"] = ht
    ex["out"] = out
    return ex


def catdisc_forward(sd: SD, x: Tensor, t: Tensor, masks=None, p: float = 0.25) -> Dict[str, Tensor]:
    """Discriminator.forward (model/GANSurv.py:61-68): out = fc(cat[EmbedX(x), time_embed(t)]). x: [N,C]; t: [1,1]."""
    ht = time_embed(sd, t)                                                 # GANSurv.py:63
    ex = embedx_forward(sd, x, masks, p)                                   # GANSurv.py:64
    ex["ht"] = ht
    ex["out"] = _lin(torch.cat([ex["hx"], ht], dim=1), sd, "fc")           # GANSurv.py:65-67
    return ex


# ----------------------------------------------------------------------------------------------
# losses (loss/utils.py)
# ----------------------------------------------------------------------------------------------
def real_fake_loss(real: Optional[Tensor], fake: Tensor, which: str = "bce") -> Tensor:
    """loss/utils.py:182-203, including the non-standard bce: -mean(1 - log(sig(fake)+1e-8)) - mean(log(sig(real)+1e-8))."""
    fake = fake.reshape(-1)
    if which == "bce":
        loss = -torch.mean(1.0 - torch.log(torch.sigmoid(fake) + 1e-8))
        if real is not None:
            loss = loss - torch.mean(torch.log(torch.sigmoid(real.reshape(-1)) + 1e-8))
    elif which == "hinge":
        loss = torch.relu(1.0 + fake).mean()
        if real is not None:
            loss = loss + torch.relu(1.0 - real.reshape(-1)).mean()
    elif which == "wasserstein":
        loss = fake.mean()
        if real is not None:
            loss = loss - real.reshape(-1).mean()
    else:
        raise ValueError(which)
    return loss


def fake_generator_loss(fake: Tensor) -> Tensor:
    """loss/utils.py:205-208."""
    return -torch.mean(fake.reshape(-1))


def recon_loss(pred_t: Tensor, t: Tensor, e: Tensor, alpha: float = 0.0, gamma: float = 0.0, norm: str = "l1") -> Tensor:
    """loss/utils.py:21-41: e|p-t| + (1-e) relu(gamma - (p - t)); l2 squares both; alpha re-weights observed part."""
    p_, t_, e_ = pred_t.reshape(-1), t.reshape(-1), e.reshape(-1)
    lo = e_ * torch.abs(p_ - t_)
    lc = (1 - e_) * torch.relu(gamma - (p_ - t_))
    if norm == "l2":
        lo, lc = lo * lo, lc * lc
    return ((1.0 - alpha) * (lo + lc) + alpha * lo).mean()


def loss_reg_l1(params: Sequence[Tensor], coef: float):
    """loss/utils.py:6-14: coef * sum |W| over ALL parameters (biases included); 0.0 if coef <= 1e-8."""
    if coef is None or coef <= 1e-8:
        return 0.0
    return coef * sum(torch.abs(w).sum() for w in params)


# ----------------------------------------------------------------------------------------------
# the adversarial step (model/model_handler.py:349-498) restated over a list of bags
# ----------------------------------------------------------------------------------------------
def disc_step_loss(sdG: SD, sdD: SD, bags: List[Tensor], ts: Tensor, es: Tensor, visible: Sequence[bool],
                   noises: List[Sequence[Optional[Tensor]]], d_masks_real=None, d_masks_fake=None,
                   which: str = "bce", noise_cfg=(0, 1), exts=None, backbone: str = "abmil") -> Dict[str, Tensor]:
    """_update_disc (model_handler.py:349-424): D.train / G.eval. Real pair only if e==1 and label visible
    (:373-377); fake pair uses pred.detach() (:396-401). Loss over all collected pairs (:412).
    d_masks_*: per-bag dropout-mask dicts for D (None = no dropout, i.e. p treated as eval)."""
    reals, fakes, preds = [], [], []
    for i, x in enumerate(bags):
        if float(es[i]) == 1.0 and visible[i]:
            mk = None if d_masks_real is None else d_masks_real[i]
            reals.append(prjdisc_forward(sdD, x, ts[i].reshape(1, 1), mk)["out"].reshape(-1))
        with torch.no_grad():
            g = generator_forward(sdG, x, noises[i], noise_cfg, None, backbone,
                                  None if exts is None else exts[i])
        pred = g["pred"].detach()
        preds.append(pred)
        mk = None if d_masks_fake is None else d_masks_fake[i]
        fakes.append(prjdisc_forward(sdD, x, pred, mk)["out"].reshape(-1))
    real = torch.cat(reals) if reals else None
    fake = torch.cat(fakes)
    return {"loss": real_fake_loss(real, fake, which), "real": real, "fake": fake, "pred": torch.cat(preds).reshape(-1)}


def gen_step_loss(sdG: SD, sdD: SD, bags: List[Tensor], ts: Tensor, es: Tensor, visible: Sequence[bool],
                  noises, g_masks=None, coef_gan: float = 0.004, coef_l1: float = 1e-5, noise_cfg=(0, 1),
                  exts=None, backbone: str = "abmil", recon_kw=None) -> Dict[str, Tensor]:
    """_update_gen (model_handler.py:426-498): D.eval / G.train. total = recon(visible bags) + coef_gan * (-mean f_fake)
    + coef_l1 * sum|W_G| (:472-485)."""
    recon_kw = recon_kw or {}
    preds, fakes = [], []
    for i, x in enumerate(bags):
        mk = None if g_masks is None else g_masks[i]
        g = generator_forward(sdG, x, noises[i], noise_cfg, mk, backbone, None if exts is None else exts[i])
        preds.append(g["pred"])
        fakes.append(prjdisc_forward(sdD, x, g["pred"], None)["out"].reshape(-1))
    fake = torch.cat(fakes)
    gen_loss = fake_generator_loss(fake)
    vis = [i for i in range(len(bags)) if visible[i]]
    if vis:
        tp = torch.cat([preds[i] for i in vis], dim=0)
        t_reg = recon_loss(tp, ts[vis], es[vis], **recon_kw)
    else:
        t_reg = torch.zeros((), dtype=fake.dtype)
    total = t_reg if coef_gan == 0.0 else t_reg + coef_gan * gen_loss
    total = total + loss_reg_l1(list(sdG.values()), coef_l1)
    return {"loss": total, "gen_loss": gen_loss, "t_reg": t_reg, "fake": fake, "pred": torch.cat(preds).reshape(-1)}


def adam_step(params: SD, grads: SD, state: Dict[str, Dict[str, Tensor]], lr: float, step: int,
              weight_decay: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.Adam as configured by the handler (model_handler.py:104-107; optim/optim_factory.py:25-37,76-77):
    L2 weight decay (added to the gradient) on tensors with ndim > 1 whose name does not end in '.bias'; none otherwise."""
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k]
        wd = weight_decay if (p.ndim > 1 and not k.endswith(".bias")) else 0.0
        if wd:
            g = g + wd * p
        st = state.setdefault(k, {"m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(st["m"], denom, value=-lr / bc1)


# ----------------------------------------------------------------------------------------------
# inference sampling + C-index (model_handler.py:598-643; eval/cindex.py)
# ----------------------------------------------------------------------------------------------
def lower_median(samples: Tensor, dim: int = 0) -> Tensor:
    """torch.median semantics used at model_handler.py:639: for an even count returns the LOWER middle value."""
    srt, _ = torch.sort(samples, dim=dim)
    k = (samples.shape[dim] - 1) // 2
    return srt.select(dim, k)


def sample_times(sdG: SD, x: Tensor, noise_list: List[Tensor], noise_cfg=(0, 1)) -> Tensor:
    """test_model inner loop (model_handler.py:624-636): G in eval mode, one prediction per noise draw. -> [S]"""
    bb = abmil_forward(sdG, x)
    out = []
    for nz in noise_list:
        noises = [nz if c == 1 else None for c in noise_cfg]
        out.append(generator_head(sdG, bb["H"], noises, noise_cfg).reshape(-1))
    return torch.cat(out)


def concordance_index(t: np.ndarray, e: np.ndarray, y_pred: np.ndarray, tied_tol: float = 1e-8) -> float:
    """Harrell's C as the reference computes it: concordance_index(y_true, y_pred) calls
    concordance_index_censored(e, t, -y_pred) (eval/cindex.py:10-40,106-200).  A pair (i,j) is comparable
    when t_i < t_j and e_i == 1; it is concordant when risk_i > risk_j (risk = -pred); |risk diff| <= tol is a
    tie counted 0.5.  Pairs with equal times are never comparable when both had events; if only one of them
    had the event, the event one is 'i' (eval/cindex.py:82-100)."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    e = np.asarray(e).reshape(-1).astype(bool)
    risk = -np.asarray(y_pred, dtype=np.float64).reshape(-1)
    num = 0.0
    den = 0.0
    for i in np.nonzero(e)[0]:
        comp = (t > t[i]) | ((t == t[i]) & (~e))
        comp[i] = False
        n = int(comp.sum())
        if n == 0:
            continue
        d = risk[comp] - risk[i]
        ties = np.abs(d) <= tied_tol
        con = (d < 0) & (~ties)
        num += con.sum() + 0.5 * ties.sum()
        den += n
    if den == 0:
        raise ValueError("no comparable pairs")
    return float(num / den)


# ----------------------------------------------------------------------------------------------
# level-2 -> level-1 patch index map (tools/big_to_small_patching.py:40-46,59-76)
# ----------------------------------------------------------------------------------------------
def region_index_map(coords_l2: np.ndarray, patch_size: int = 256, scale: int = 4) -> np.ndarray:
    """coords_x5_to_x20: parent k (file order) -> 16 children at rows 16k + 4j + i with coords
    c_k + (i*psize, j*psize), j outer / i inner (get_scaled_matrix :40-46); output float64 (:60,71-72)."""
    coords_l2 = np.asarray(coords_l2)
    M = coords_l2.shape[0]
    out = np.zeros((M * scale * scale, 2), dtype=np.float64)
    for k in range(M):
        for j in range(scale):
            for i in range(scale):
                out[k * scale * scale + j * scale + i, 0] = coords_l2[k, 0] + i * patch_size
                out[k * scale * scale + j * scale + i, 1] = coords_l2[k, 1] + j * patch_size
    return out


def mask_regions_zero(bag: Tensor, keep_regions: np.ndarray, scale: int = 4) -> Tensor:
    """random_mask_square_instance(..., 'mask_zero') given the kept region ids (utils/func.py:14-40)."""
    out = torch.zeros_like(bag)
    s2 = scale * scale
    for r in np.sort(np.asarray(keep_regions)):
        out[r * s2:(r + 1) * s2] = bag[r * s2:(r + 1) * s2]
    return out


# ----------------------------------------------------------------------------------------------
# deterministic synthetic inputs / parameters shared by oracle, goldens, tests and bench
# ----------------------------------------------------------------------------------------------
G_SHAPES = lambda C=1024, h=384, o=384: {  # noqa: E731  (ABMIL generator, noise 0-1, hops 1)
    "MLPs.0.0.weight": (o // 2, o), "MLPs.0.0.bias": (o // 2,),
    "MLPs.1.0.weight": (1, o), "MLPs.1.0.bias": (1,),
    "backbone.attention_net.0.weight": (h, C), "backbone.attention_net.0.bias": (h,),
    "backbone.attention_net.3.attention_a.0.weight": (h, h), "backbone.attention_net.3.attention_a.0.bias": (h,),
    "backbone.attention_net.3.attention_b.0.weight": (h, h), "backbone.attention_net.3.attention_b.0.bias": (h,),
    "backbone.attention_net.3.attention_c.weight": (1, h), "backbone.attention_net.3.attention_c.bias": (1,),
    "backbone.rho.0.weight": (o, h), "backbone.rho.0.bias": (o,),
}

G_CLUSTER_SHAPES = lambda C=1024, h=384: {  # noqa: E731  DeepAttMISL generator
    "MLPs.0.0.weight": (h // 2, h), "MLPs.0.0.bias": (h // 2,),
    "MLPs.1.0.weight": (1, h), "MLPs.1.0.bias": (1,),
    "backbone.phis.0.weight": (h, C, 1, 1), "backbone.phis.0.bias": (h,),
    "backbone.attention_net.0.weight": (h, h), "backbone.attention_net.0.bias": (h,),
    "backbone.attention_net.3.attention_a.0.weight": (h, h), "backbone.attention_net.3.attention_a.0.bias": (h,),
    "backbone.attention_net.3.attention_b.0.weight": (h, h), "backbone.attention_net.3.attention_b.0.bias": (h,),
    "backbone.attention_net.3.attention_c.weight": (1, h), "backbone.attention_net.3.attention_c.bias": (1,),
}

def G_ESAT_SHAPES(C=1024, d=384):
    """State dict of Generator(backbone=DualTrans_HS) as printed from the reference modules (load_backbone('patch'))."""
    L = "backbone.patch_encoder_layer.layers.0."
    sh = {"MLPs.0.0.weight": (d // 2, d), "MLPs.0.0.bias": (d // 2,), "MLPs.1.0.weight": (1, d), "MLPs.1.0.bias": (1,),
          "backbone.patch_embedding_layer.conv.weight": (d, C, 1, 1), "backbone.patch_embedding_layer.conv.bias": (d,),
          "backbone.patch_embedding_layer.norm.weight": (d,), "backbone.patch_embedding_layer.norm.bias": (d,),
          L + "self_attn.in_proj_weight": (3 * d, d), L + "self_attn.in_proj_bias": (3 * d,),
          L + "self_attn.out_proj.weight": (d, d), L + "self_attn.out_proj.bias": (d,),
          L + "linear1.weight": (d, d), L + "linear1.bias": (d,), L + "linear2.weight": (d, d), L + "linear2.bias": (d,),
          L + "norm1.weight": (d,), L + "norm1.bias": (d,), L + "norm2.weight": (d,), L + "norm2.bias": (d,),
          "backbone.pool.fc1.0.weight": (d, d), "backbone.pool.fc1.0.bias": (d,),
          "backbone.pool.score.0.weight": (d, d), "backbone.pool.score.0.bias": (d,),
          "backbone.pool.fc2.weight": (1, d), "backbone.pool.fc2.bias": (1,)}
    return sh


def DCAT_SHAPES(C=1024, d=128, ty=(64, 128)):
    """State dict of the concat Discriminator (model/GANSurv.py:52-60): PrjDiscriminator's minus prj_layer, plus fc."""
    sh = {k: v for k, v in D_SHAPES(C, d, ty).items() if not k.startswith("prj_layer")}
    sh["fc.weight"], sh["fc.bias"] = (1, d + ty[1]), (1,)
    return sh


D_SHAPES = lambda C=1024, d=128, ty=(64, 128): {  # noqa: E731
    "net_pair_one.embedding.conv.weight": (d, C, 1, 1), "net_pair_one.embedding.conv.bias": (d,),
    "net_pair_one.embedding.norm.weight": (d,), "net_pair_one.embedding.norm.bias": (d,),
    "net_pair_one.fc1.0.weight": (d // 2, d), "net_pair_one.fc1.0.bias": (d // 2,),
    "net_pair_one.fc1.3.weight": (d, d // 2), "net_pair_one.fc1.3.bias": (d,),
    "net_pair_one.pool.fc1.0.weight": (d, d), "net_pair_one.pool.fc1.0.bias": (d,),
    "net_pair_one.pool.score.0.weight": (d, d), "net_pair_one.pool.score.0.bias": (d,),
    "net_pair_one.pool.fc2.weight": (1, d), "net_pair_one.pool.fc2.bias": (1,),
    "net_pair_one.fc2.0.weight": (d // 2, d), "net_pair_one.fc2.0.bias": (d // 2,),
    "net_pair_one.fc2.3.weight": (d, d // 2), "net_pair_one.fc2.3.bias": (d,),
    "net_pair_two.0.0.weight": (ty[0], 1), "net_pair_two.0.0.bias": (ty[0],),
    "net_pair_two.1.0.weight": (ty[1], ty[0]), "net_pair_two.1.0.bias": (ty[1],),
    "prj_layer.weight": (1, d), "prj_layer.bias": (1,),
}


def synth_state_dict(shapes: Dict[str, tuple], seed: int, dtype=torch.float32) -> SD:
    """Deterministic parameters independent of torch's RNG/init order: for each tensor (sorted by name)
    uniform(-b, b) with b = 1/sqrt(fan_in) (weights) or 0.1 (1-D); LayerNorm weight is shifted to ~1."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) > 1:
            fan_in = int(np.prod(shp[1:]))
            b = 1.0 / math.sqrt(fan_in)
        else:
            b = 0.1
        v = rng.uniform(-b, b, size=shp)
        if name.endswith(("norm.weight", "norm1.weight", "norm2.weight")):
            v = v + 1.0
        sd[name] = torch.tensor(v, dtype=dtype)
    return sd


def synth_bag(n_rows: int, seed: int, C: int = 1024, nonneg: bool = False, dtype=torch.float32) -> Tensor:
    """Synthetic bag: randn features (precedent model_stats.py:93) or relu(randn)*0.5."""
    rng = np.random.default_rng(1000 + seed)
    v = rng.standard_normal((n_rows, C)).astype(np.float32)
    if nonneg:
        v = np.maximum(v, 0) * 0.5
    return torch.tensor(v, dtype=dtype)


def synth_labels(n_bags: int, seed: int, event_rate: float = 0.347):
    """t ~ U(0,1), e ~ Bernoulli(0.347) (NLST event rate)."""
    rng = np.random.default_rng(2000 + seed)
    t = torch.tensor(rng.uniform(0.02, 0.98, size=n_bags), dtype=torch.float32)
    e = torch.tensor((rng.uniform(size=n_bags) < event_rate).astype(np.float32))
    return t, e


def synth_masks(shape, keep_prob: float, seed: int) -> Tensor:
    rng = np.random.default_rng(3000 + seed)
    return torch.tensor((rng.uniform(size=shape) < keep_prob).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# one full adversarial step on CPU (used as the timed CPU baseline and by the step parity tests)
# ----------------------------------------------------------------------------------------------
class CpuTrainer:
    """Restates MyHandler's optimiser setup (model/model_handler.py:104-107; optim/optim_factory.py:25-37,76-77) and one
    `_update_disc` + `_update_gen` round (:349-498) on top of the functional forward above, with torch.optim.Adam —
    the same ATen kernels the reference's modules dispatch to on CPU."""

    def __init__(self, sdG: SD, sdD: SD, lr: float = 8e-5, wd_g: float = 5e-4, coef_gan: float = 0.004,
                 coef_l1: float = 1e-5, which: str = "bce", backbone: str = "abmil"):
        self.backbone = backbone
        self.sdG = {k: v.clone().requires_grad_(True) for k, v in sdG.items()}
        self.sdD = {k: v.clone().requires_grad_(True) for k, v in sdD.items()}
        no_decay = [v for k, v in self.sdG.items() if v.ndim == 1 or k.endswith(".bias")]
        decay = [v for k, v in self.sdG.items() if not (v.ndim == 1 or k.endswith(".bias"))]
        self.optG = torch.optim.Adam([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd_g}], lr=lr)
        self.optD = torch.optim.Adam(list(self.sdD.values()), lr=lr, betas=(0.9, 0.999), weight_decay=0.0)
        self.coef_gan, self.coef_l1, self.which = coef_gan, coef_l1, which

    def step(self, bags, ts, es, visible, noise_d, noise_g, d_masks_real=None, d_masks_fake=None, g_masks=None, exts=None):
        nzd = [[None, n.reshape(1, -1)] for n in noise_d]
        nzg = [[None, n.reshape(1, -1)] for n in noise_g]
        d = disc_step_loss(self.sdG, self.sdD, bags, ts, es, visible, nzd, d_masks_real, d_masks_fake, self.which,
                           exts=exts, backbone=self.backbone)
        self.optD.zero_grad()
        d["loss"].backward()
        d_grads = {k: v.grad.clone() for k, v in self.sdD.items()}
        self.optD.step()
        g = gen_step_loss(self.sdG, self.sdD, bags, ts, es, visible, nzg, g_masks, self.coef_gan, self.coef_l1,
                          exts=exts, backbone=self.backbone)
        self.optG.zero_grad()
        for v in self.sdD.values():
            v.grad = None
        g["loss"].backward()
        g_grads = {k: v.grad.clone() for k, v in self.sdG.items()}
        self.optG.step()
        return {"dis_loss": float(d["loss"]), "gen_loss": float(g["gen_loss"]), "t_reg": float(g["t_reg"]),
                "total": float(g["loss"]), "pred_d": d["pred"].detach(), "pred_g": g["pred"].detach(),
                "fake_d": d["fake"].detach(), "fake_g": g["fake"].detach(), "d_grads": d_grads, "g_grads": g_grads}


def random_g_masks(n_rows: int, h: int, o: int, gen: torch.Generator):
    """Bernoulli keep masks for the five generator dropout sites (p = 0.25 backbone, 0.6 head)."""
    r = lambda shp, keep: (torch.rand(shp, generator=gen) < keep).float()  # noqa: E731
    return {"h": r((n_rows, h), 0.75), "a": r((n_rows, h), 0.75), "b": r((n_rows, h), 0.75), "rho": r((1, o), 0.75),
            "mlp0": r((1, o // 2), 0.4)}


def random_d_masks(n_regions: int, d: int, gen: torch.Generator):
    r = lambda shp: (torch.rand(shp, generator=gen) < 0.75).float()  # noqa: E731
    return {"fc1": r((n_regions, d // 2)), "ga": r((n_regions, d)), "gs": r((n_regions, d)), "fc2": r((1, d // 2))}
