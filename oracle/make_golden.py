"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (build container only).

    python oracle/make_golden.py            # writes tests/golden/

TEST INFRASTRUCTURE.  The reference has no golden vectors of its own, so the pin
is the live reference: its nn.Modules are built exactly as model/model_handler.py:74-87 builds them,
loaded with deterministic synthetic parameters (oracle.advmil_oracle.synth_state_dict), fed
deterministic synthetic bags, and their outputs / autograd gradients are stored (sub-sampled to keep
fixtures small).  Dropout sites are made reproducible by swapping the reference's nn.Dropout
*instances* (not its source) for a fixed-mask module.  tests/test_oracle_golden.py then checks the
oracle restatement against these files; the GPU tests check the CUDA path against oracle + goldens.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import advmil_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class FixedDropout(nn.Module):
    """Stands in for an nn.Dropout instance inside the reference model: y = x * mask / (1-p)."""

    def __init__(self, p):
        super().__init__()
        self.p = p
        self.mask = None

    def forward(self, x):
        if not self.training or self.mask is None:
            return x
        return x * self.mask.to(x.dtype).reshape(x.shape) / (1.0 - self.p)


def sub(v, n=512):
    v = v.detach().reshape(-1).double().numpy()
    stride = max(1, v.size // n)
    return v[::stride].copy()


def build_ref_G(ref, dims=(1024, 384, 384), mode="abmil", gen_dropout=0.6):
    backbone = ref.backbone.load_backbone(mode, list(dims))
    args_noise = SimpleNamespace(noise=[0, 1], hops=1, noise_dist="uniform")
    return ref.GANSurv.Generator(dims[2], 1, backbone, args_noise, False, gen_dropout, "sigmoid")


def build_ref_D(ref, C=1024, d=128, iprd="instance", prj="x"):
    ax = SimpleNamespace(in_dim=C, out_dim=d, ksize=1, backbone="avgpool", dropout=0.25)
    ay = SimpleNamespace(in_dim=1, hid_dims=[64, 128] if d == 128 else [d // 2, d], norm=False, dropout=0.0)
    return ref.GANSurv.PrjDiscriminator(ax, ay, prj_path=prj, inner_product=iprd)


def swap_dropouts_G(G):
    d = {}
    bb = G.backbone
    bb.attention_net[2] = d["h"] = FixedDropout(0.25)
    bb.attention_net[3].attention_a[2] = d["a"] = FixedDropout(0.25)
    bb.attention_net[3].attention_b[2] = d["b"] = FixedDropout(0.25)
    if hasattr(bb, "rho"):
        bb.rho[2] = d["rho"] = FixedDropout(0.25)
    G.MLPs[0][2] = d["mlp0"] = FixedDropout(G.MLPs[0][2].p)
    return d


def swap_dropouts_D(D):
    d = {}
    e = D.net_pair_one
    e.fc1[2] = d["fc1"] = FixedDropout(0.25)
    e.pool.fc1[2] = d["ga"] = FixedDropout(0.25)
    e.pool.score[2] = d["gs"] = FixedDropout(0.25)
    e.fc2[2] = d["fc2"] = FixedDropout(0.25)
    return d


def g_masks(N, h, o, seed):
    return {"h": O.synth_masks((N, h), 0.75, seed), "a": O.synth_masks((N, h), 0.75, seed + 1),
            "b": O.synth_masks((N, h), 0.75, seed + 2), "rho": O.synth_masks((1, o), 0.75, seed + 3),
            "mlp0": O.synth_masks((1, o // 2), 0.4, seed + 4)}


def d_masks(R, d, seed):
    return {"fc1": O.synth_masks((R, d // 2), 0.75, seed), "ga": O.synth_masks((R, d), 0.75, seed + 1),
            "gs": O.synth_masks((R, d), 0.75, seed + 2), "fc2": O.synth_masks((1, d // 2), 0.75, seed + 3)}


def run_ref_G(ref, G, x, noise, zero_noise=False):
    """Generator.forward draws noise from the CPU generator (utils/func.py:154-164); to feed a chosen
    noise tensor, the module-level generate_noise symbol that GANSurv.py imported is patched for the call."""
    orig = ref.GANSurv.generate_noise
    ref.GANSurv.generate_noise = lambda *dims, to_device="cpu", distribution="uniform": noise.clone()
    try:
        return G(x.unsqueeze(0), torch.Tensor([0]).unsqueeze(0), zero_noise=zero_noise)
    finally:
        ref.GANSurv.generate_noise = orig


def case_generator(ref, name, dims, N, train, seed, nonneg=False):
    C, h, o = dims
    G = build_ref_G(ref, dims)
    sd = O.synth_state_dict(O.G_SHAPES(C, h, o), seed)
    G.load_state_dict(sd)
    x = O.synth_bag(N, seed, C, nonneg)
    noise = torch.tensor(np.random.default_rng(seed + 7).uniform(size=(1, o // 2)), dtype=torch.float32)
    drops = swap_dropouts_G(G)
    masks = None
    if train:
        G.train()
        masks = g_masks(N, h, o, seed * 10)
        for k, m in drops.items():
            m.mask = masks[k]
    else:
        G.eval()
    # hooks for intermediates
    inter = {}
    def _hook_attn(m, i, out):
        inter["s"] = out[0]

    def _hook_rho(m, i, out):
        inter["z"], inter["H"] = i[0], out

    G.backbone.attention_net[3].register_forward_hook(_hook_attn)
    G.backbone.rho.register_forward_hook(_hook_rho)
    pred = run_ref_G(ref, G, x, noise)
    G.zero_grad()
    pred.sum().backward()
    out = {"pred": pred.detach().double().numpy(), "s": sub(inter["s"]), "z": inter["z"].detach().double().numpy(),
           "H": inter["H"].detach().double().numpy(),
           "cfg": np.array([C, h, o, N, int(train), seed, int(nonneg)])}
    for k, p in G.named_parameters():
        out["grad." + k] = sub(p.grad)
        out["gsum." + k] = np.array(p.grad.double().sum().item())
    # oracle cross-check at generation time
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    og = O.generator_forward(sdr, x, [None, noise], (0, 1), masks)
    og["pred"].sum().backward()
    err = float((og["pred"].detach() - pred.detach()).abs().max())
    gerr = max(float((sdr[k].grad - p.grad).abs().max() / (p.grad.abs().max() + 1e-3)) for k, p in G.named_parameters())
    print(f"[golden] {name}: pred {pred.item():.8f} oracle|d|={err:.2e} grad rel err={gerr:.2e}")
    assert err < 1e-6 and gerr < 1e-4
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_disc(ref, name, C, d, N, train, seed, iprd="instance", prj="x"):
    D = build_ref_D(ref, C, d, iprd, prj)
    ty = (64, 128) if d == 128 else (d // 2, d)
    sd = O.synth_state_dict(O.D_SHAPES(C, d, ty), seed + 50)
    D.load_state_dict(sd)
    x = O.synth_bag(N, seed, C)
    t = torch.tensor([[0.37]], dtype=torch.float32, requires_grad=True)
    drops = swap_dropouts_D(D)
    masks = None
    if train:
        D.train()
        masks = d_masks(N // 16, d, seed * 10 + 5)
        for k, m in drops.items():
            m.mask = masks[k]
    else:
        D.eval()
    inter = {}
    def _mk(key, pick=None):
        def _h(m, i, o_):
            inter[key] = o_ if pick is None else pick(o_)
        return _h

    D.net_pair_one.embedding.register_forward_hook(_mk("emb"))
    D.net_pair_one.fc1.register_forward_hook(_mk("fi"))
    D.net_pair_one.register_forward_hook(_mk("hx", lambda o_: o_[0] if isinstance(o_, tuple) else o_))
    out_t = D(x.unsqueeze(0), t)
    D.zero_grad()
    out_t.sum().backward()
    out = {"out": out_t.detach().double().numpy(), "emb": sub(inter["emb"], 2048), "fi": sub(inter["fi"], 2048),
           "hx": inter["hx"].detach().double().numpy(), "dt": t.grad.double().numpy(),
           "cfg": np.array([C, d, N, int(train), seed, int(iprd == "instance"), int(prj == "x")])}
    for k, p in D.named_parameters():
        out["grad." + k] = sub(p.grad)
        out["gsum." + k] = np.array(p.grad.double().sum().item())
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t2 = t.detach().clone().requires_grad_(True)
    od = O.prjdisc_forward(sdr, x, t2, masks, iprd, prj)
    od["out"].sum().backward()
    err = float((od["out"].detach() - out_t.detach()).abs().max())
    gerr = max(float((sdr[k].grad - p.grad).abs().max() / (p.grad.abs().max() + 1e-3)) for k, p in D.named_parameters())
    print(f"[golden] {name}: out {out_t.item():.8f} oracle|d|={err:.2e} grad rel err={gerr:.2e} dt={t.grad.item():.6e}")
    assert err < 1e-6 and gerr < 1e-4
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_disc_cat(ref, name, C, d, N, train, seed):
    """The concat Discriminator (model/GANSurv.py:52-68) built like model_handler.py:83-86 builds it.  The embedding
    gradients of this variant flow only through the attention pooling and are ill-conditioned in fp32 (the reference's
    own fp32 conv2d path is off by ~2e-2 of the tensor maximum from its float64 evaluation), so the stored gradients come
    from the reference modules run in float64; the output is the fp32 one."""
    ax = SimpleNamespace(in_dim=C, out_dim=d, ksize=1, backbone="avgpool", dropout=0.25)
    ty = (64, 128) if d == 128 else (d // 2, d)
    ay = SimpleNamespace(in_dim=1, hid_dims=list(ty), norm=False, dropout=0.0)
    sd = O.synth_state_dict(O.DCAT_SHAPES(C, d, ty), seed + 50)
    x = O.synth_bag(N, seed, C)
    masks = d_masks(N // 16, d, seed * 10 + 5) if train else None
    runs = {}
    for dt in (torch.float32, torch.float64):
        D = ref.GANSurv.Discriminator(ax, ay).to(dt)
        D.load_state_dict({k: v.to(dt) for k, v in sd.items()})
        t = torch.tensor([[0.61]], dtype=dt, requires_grad=True)
        drops = swap_dropouts_D(D)
        D.train(bool(train))
        if train:
            for k, m in drops.items():
                m.mask = masks[k]
        out_t = D(x.to(dt).unsqueeze(0), t)
        D.zero_grad()
        out_t.sum().backward()
        runs[dt] = (out_t.detach(), t.grad.detach(), {k: p.grad.detach() for k, p in D.named_parameters()})
    out32, _, _ = runs[torch.float32]
    _, dt64, g64 = runs[torch.float64]
    out = {"out": out32.double().numpy(), "dt": dt64.double().numpy(), "cfg": np.array([C, d, N, int(train), seed])}
    for k, g in g64.items():
        out["grad." + k] = sub(g)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t2 = torch.tensor([[0.61]], dtype=torch.float32, requires_grad=True)
    od = O.catdisc_forward(sdr, x, t2, masks)
    od["out"].sum().backward()
    err = float((od["out"].detach() - out32).abs().max())
    gerr = max(float((sdr[k].grad.double() - g).abs().max() / (g.abs().max() + 1e-6)) for k, g in g64.items()
               if float(g.abs().max()) > 1e-8)       # pool.fc2.bias: mathematically zero (softmax shift invariance)
    print(f"[golden] {name}: out {float(out32):.8f} oracle|d|={err:.2e} grad rel err vs float64 reference={gerr:.2e}")
    assert err < 1e-6 and gerr < 1e-4
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_cluster(ref, name, dims, N, seed, empty_cluster):
    C, h, _ = dims
    G = build_ref_G(ref, dims, mode="cluster")
    sd = O.synth_state_dict(O.G_CLUSTER_SHAPES(C, h), seed + 20)
    G.load_state_dict(sd)
    G.eval()
    x = O.synth_bag(N, seed, C)
    rng = np.random.default_rng(seed + 21)
    cid = rng.integers(0, 8, size=N)
    if empty_cluster:
        cid[cid == 5] = 2
    cid_t = torch.tensor(cid, dtype=torch.float32)
    noise = torch.tensor(np.random.default_rng(seed + 7).uniform(size=(1, h // 2)), dtype=torch.float32)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # backbone.py:115 hard-codes .cuda() for empty clusters
    orig = ref.GANSurv.generate_noise
    ref.GANSurv.generate_noise = lambda *dims_, to_device="cpu", distribution="uniform": noise.clone()
    ref_raised = False
    try:
        pred = G(x.unsqueeze(0), cid_t)       # model_stats.py:134 passes [N] ids
    except RuntimeError as ex:
        # torch >= 2.x: conv2d rejects the zero-width input of an empty cluster before the reference's own
        # zeros fallback (backbone.py:114-115) is reached.  The intended semantics (zeros for an empty cluster)
        # are what the oracle implements; this fixture then records the oracle's output, flagged ref_raised.
        ref_raised = True
        print(f"[golden] {name}: reference raised on empty cluster ({str(ex)[:60]}...) -> oracle-only fixture")
    finally:
        torch.Tensor.cuda = orig_cuda
        ref.GANSurv.generate_noise = orig
    if ref_raised:
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        og = O.generator_forward(sdr, x, [None, noise], (0, 1), None, "cluster", cid_t)
        og["pred"].sum().backward()
        out = {"pred": og["pred"].detach().double().numpy(), "cid": cid.astype(np.int32), "ref_raised": np.array(1),
               "cfg": np.array([C, h, N, seed, int(empty_cluster)])}
        for k in sd:
            out["grad." + k] = sub(sdr[k].grad)
            out["gsum." + k] = np.array(sdr[k].grad.double().sum().item())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        return
    G.zero_grad()
    pred.sum().backward()
    out = {"pred": pred.detach().double().numpy(), "cid": cid.astype(np.int32),
           "cfg": np.array([C, h, N, seed, int(empty_cluster)])}
    for k, p in G.named_parameters():
        out["grad." + k] = sub(p.grad)
        out["gsum." + k] = np.array(p.grad.double().sum().item())
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    og = O.generator_forward(sdr, x, [None, noise], (0, 1), None, "cluster", cid_t)
    og["pred"].sum().backward()
    err = float((og["pred"].detach() - pred.detach()).abs().max())
    gerr = max(float((sdr[k].grad - p.grad).abs().max() / (p.grad.abs().max() + 1e-3)) for k, p in G.named_parameters())
    print(f"[golden] {name}: pred {pred.item():.8f} oracle|d|={err:.2e} grad rel err={gerr:.2e}")
    assert err < 1e-6 and gerr < 1e-4
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def esat_masks(R, d, seed, nhead=8):
    return {"attn": O.synth_masks((nhead, R, R), 0.75, seed), "sa": O.synth_masks((R, d), 0.75, seed + 1),
            "ff1": O.synth_masks((R, d), 0.75, seed + 2), "ff2": O.synth_masks((R, d), 0.75, seed + 3),
            "ga": O.synth_masks((R, d), 0.75, seed + 4), "gs": O.synth_masks((R, d), 0.75, seed + 5),
            "mlp0": O.synth_masks((1, d // 2), 0.4, seed + 6)}


def case_esat(ref, name, C, d, N, train, seed, with_coord):
    """Generator(backbone=DualTrans_HS) as load_backbone('patch') builds it (model/backbone.py:31-35,171-196).  Train
    mode: the nn.Dropout instances are swapped for fixed-mask modules; the dropout on the attention probabilities lives
    inside torch's own MultiheadAttention code, so the layer's self_attn is called with need_weights=True (torch's math
    path, same arithmetic) and torch.nn.functional.dropout is patched for the [nhead, R, R] tensor only."""
    import torch.nn.functional as F
    G = build_ref_G(ref, (C, d, d), mode="patch")
    sd = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), seed)
    G.load_state_dict(sd)
    x = O.synth_bag(N, seed, C)
    R = N // 16
    noise = torch.tensor(np.random.default_rng(seed + 7).uniform(size=(1, d // 2)), dtype=torch.float32)
    coord = torch.tensor(np.random.default_rng(seed + 8).integers(0, 60, size=(R, 2)), dtype=torch.int64) if with_coord else None
    layer = G.backbone.patch_encoder_layer.layers[0]
    drops = {}
    layer.dropout1 = drops["sa"] = FixedDropout(0.25)
    layer.dropout = drops["ff1"] = FixedDropout(0.25)
    layer.dropout2 = drops["ff2"] = FixedDropout(0.25)
    G.backbone.pool.fc1[2] = drops["ga"] = FixedDropout(0.25)
    G.backbone.pool.score[2] = drops["gs"] = FixedDropout(0.25)
    G.MLPs[0][2] = drops["mlp0"] = FixedDropout(G.MLPs[0][2].p)
    masks = None
    orig_dropout, orig_sa = F.dropout, layer.self_attn.forward
    if train:
        G.train()
        masks = esat_masks(R, d, seed * 10)
        for k, m in drops.items():
            m.mask = masks[k]
        layer.self_attn.forward = lambda *a, **k: orig_sa(*a, **{**k, "need_weights": True})

        def fixed_dropout(inp, p=0.5, training=True, inplace=False):
            if training and tuple(inp.shape) == tuple(masks["attn"].shape):
                return inp * masks["attn"].to(inp.dtype) / (1.0 - p)
            return orig_dropout(inp, p, training, inplace)
        F.dropout = fixed_dropout
    else:
        G.eval()
    inter = {}
    G.backbone.patch_embedding_layer.register_forward_hook(lambda m, i, out: inter.__setitem__("emb", out))
    G.backbone.patch_encoder_layer.register_forward_hook(lambda m, i, out: inter.__setitem__("x2", out))
    G.backbone.register_forward_hook(lambda m, i, out: inter.__setitem__("H", out))
    orig = ref.GANSurv.generate_noise
    ref.GANSurv.generate_noise = lambda *dims, to_device="cpu", distribution="uniform": noise.clone()
    try:
        pred = G(x.unsqueeze(0), None if coord is None else coord.unsqueeze(0))
    finally:
        ref.GANSurv.generate_noise = orig
        F.dropout = orig_dropout
    G.zero_grad()
    pred.sum().backward()
    out = {"pred": pred.detach().double().numpy(), "H": inter["H"].detach().double().numpy(), "emb": sub(inter["emb"]),
           "x2": sub(inter["x2"]), "cfg": np.array([C, d, N, int(train), seed, int(with_coord)])}
    if coord is not None:
        out["coord"] = coord.numpy()
    for k, p in G.named_parameters():
        out["grad." + k] = sub(p.grad)
        out["gsum." + k] = np.array(p.grad.double().sum().item())
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    og = O.generator_forward(sdr, x, [None, noise], (0, 1), masks, backbone="patch", coord=coord)
    og["pred"].sum().backward()
    err = float((og["pred"].detach() - pred.detach()).abs().max())
    gerr = max(float((sdr[k].grad - p.grad).abs().max() / (p.grad.abs().max() + 1e-3)) for k, p in G.named_parameters())
    print(f"[golden] {name}: pred {pred.item():.8f} oracle|d|={err:.2e} grad rel err={gerr:.2e}")
    assert err < 1e-6 and gerr < 1e-4
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_step(ref, name, dims, d, Ns, seed, n_steps=2):
    """Full adversarial step(s) with the reference modules, losses and optimisers, restating the loop of
    model_handler.py:349-498 around them (the handler itself hard-codes .cuda() and wandb)."""
    C, h, o = dims
    G = build_ref_G(ref, dims)
    D = build_ref_D(ref, C, d)
    ty = (64, 128) if d == 128 else (d // 2, d)
    G.load_state_dict(O.synth_state_dict(O.G_SHAPES(C, h, o), seed))
    D.load_state_dict(O.synth_state_dict(O.D_SHAPES(C, d, ty), seed + 50))
    gd, dd = swap_dropouts_G(G), swap_dropouts_D(D)
    cfgopt = SimpleNamespace(opt="adam", weight_decay=5e-4, lr=8e-5, opt_eps=None, opt_betas=None, momentum=None)
    optG = ref.optim.create_optimizer(cfgopt, G)
    optD = torch.optim.Adam(D.parameters(), lr=8e-5, betas=(0.9, 0.999), weight_decay=0.0)
    l1 = ref.loss.loss_reg_l1(1e-5)
    B = len(Ns)
    bags = [O.synth_bag(n, seed + i, C) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, seed)
    es[0] = 1.0
    visible = [(i % 3) != 2 for i in range(B)]
    out = {"cfg": np.array([C, h, o, d, seed, n_steps] + list(Ns)), "t": ts.numpy(), "e": es.numpy(),
           "visible": np.array(visible)}
    for step in range(n_steps):
        rng = np.random.default_rng(seed + 100 * step)
        nzD = [torch.tensor(rng.uniform(size=(1, o // 2)), dtype=torch.float32) for _ in range(B)]
        nzG = [torch.tensor(rng.uniform(size=(1, o // 2)), dtype=torch.float32) for _ in range(B)]
        # ---- D step (model_handler.py:349-424)
        D.train(); G.eval()
        reals, fakes, preds = [], [], []
        for i in range(B):
            x = bags[i].unsqueeze(0)
            t = ts[i].reshape(1, 1)
            if es[i] == 1 and visible[i]:
                mk = d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i)
                for k, m in dd.items():
                    m.mask = mk[k]
                reals.append(D(x, t).view(-1))
            pred = run_ref_G(ref, G, bags[i], nzD[i])
            preds.append(pred)
            mk = d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i + 5)
            for k, m in dd.items():
                m.mask = mk[k]
            fakes.append(D(x, pred.detach()).view(-1))
        optD.zero_grad()
        dis_loss = ref.loss.real_fake_loss(torch.cat(reals) if reals else None, torch.cat(fakes), which="bce")
        dis_loss.backward()
        if step == 0:
            for k, p in D.named_parameters():
                out["dgrad." + k] = sub(p.grad)
        optD.step()
        # ---- G step (model_handler.py:426-498)
        D.eval(); G.train()
        preds2, ff = [], []
        for i in range(B):
            mk = g_masks(Ns[i], h, o, seed + 2000 * step + 10 * i)
            for k, m in gd.items():
                m.mask = mk[k]
            pred = run_ref_G(ref, G, bags[i], nzG[i])
            preds2.append(pred)
            ff.append(D(bags[i].unsqueeze(0), pred).view(-1))
        optG.zero_grad()
        gen_loss = ref.loss.fake_generator_loss(torch.cat(ff))
        vis = [i for i in range(B) if visible[i]]
        t_reg = ref.loss.recon_loss(torch.cat([preds2[i] for i in vis]), ts[vis].reshape(-1, 1), es[vis].reshape(-1, 1),
                                    alpha=0.0, gamma=0.0, norm="l1")
        total = t_reg + 0.004 * gen_loss
        total = total + l1(G.parameters())
        total.backward()
        if step == 0:
            for k, p in G.named_parameters():
                out["ggrad." + k] = sub(p.grad)
        optG.step()
        out[f"dis_loss{step}"] = np.array(dis_loss.item())
        out[f"gen_loss{step}"] = np.array(gen_loss.item())
        out[f"t_reg{step}"] = np.array(t_reg.item())
        out[f"total{step}"] = np.array(total.item())
        out[f"pred_d{step}"] = torch.cat(preds).detach().reshape(-1).double().numpy()
        out[f"pred_g{step}"] = torch.cat(preds2).detach().reshape(-1).double().numpy()
        out[f"fake_d{step}"] = torch.cat(fakes).detach().double().numpy()
        out[f"fake_g{step}"] = torch.cat(ff).detach().double().numpy()
        print(f"[golden] {name} step{step}: dis {dis_loss.item():.6f} gen {gen_loss.item():.6f} "
              f"t_reg {t_reg.item():.6f} total {total.item():.6f}")
    for k, p in G.named_parameters():
        out["gparam." + k] = sub(p)
    for k, p in D.named_parameters():
        out["dparam." + k] = sub(p)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_misc(ref):
    out = {}
    # (1) region <-> patch index map: tools/big_to_small_patching.py get_scaled_matrix + the loop at :67-72
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2s", os.path.join(os.environ.get("ADVMIL_REFERENCE", "/root/reference"),
                                                                      "tools", "big_to_small_patching.py"))
    b2s = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b2s)
    rng = np.random.default_rng(5)
    coords = (rng.integers(0, 200, size=(37, 2)) * 1024).astype(np.int64)
    mat = b2s.get_scaled_matrix(256, 256, 4)
    scaled = np.zeros((1, 2), dtype=np.int32)
    for c in coords:
        scaled = np.concatenate((scaled, mat + c), axis=0)
    out["l2_coords"] = coords
    out["l1_coords"] = scaled[1:]
    # (2) sequence2square: row n of the sequence lands in region n // 16, position ((n%16)//4, n%4)
    seq = torch.arange(64 * 3, dtype=torch.float32).reshape(1, 64, 3)
    sq, L = ref.backbone_utils.sequence2square(seq, 4)
    out["seq2sq"] = sq.numpy()
    out["sq2seq"] = ref.backbone_utils.square2sequence(sq, L).numpy()
    # (3) C-index incl. ties in time and prediction
    n = 447
    t, e = O.synth_labels(n, 9)
    t = torch.round(t * 50) / 50
    pred = torch.tensor(np.round(np.random.default_rng(10).uniform(size=n), 2), dtype=torch.float32)
    y_true = torch.stack([t, e], dim=1).numpy()
    out["ci_t"], out["ci_e"], out["ci_pred"] = t.numpy(), e.numpy(), pred.numpy()
    out["ci"] = np.array(ref.cindex.concordance_index(y_true, pred.unsqueeze(-1).numpy()))
    # (4) lower median of 30 samples
    s = torch.tensor(np.random.default_rng(11).uniform(size=(30, 4, 1)), dtype=torch.float32)
    out["med_in"] = s.numpy()
    out["med_out"] = torch.median(s, dim=0)[0].numpy()
    # (5) generate_noise == torch.rand / randn on the CPU stream
    torch.manual_seed(123)
    a = ref.func.generate_noise(1, 192, distribution="uniform")
    b = ref.func.generate_noise(1, 192, distribution="gaussian")
    torch.manual_seed(123)
    a2 = torch.rand(1, 192)
    b2 = torch.randn(1, 192)
    out["noise_eq"] = np.array([bool(torch.equal(a, a2)), bool(torch.equal(b, b2))])
    # (6) losses on fixed scores
    fr = torch.tensor([0.3, -1.2, 2.0]); ff = torch.tensor([-0.5, 0.1, 0.7, -2.0])
    for w in ("bce", "hinge", "wasserstein"):
        out["rf_" + w] = np.array(ref.loss.real_fake_loss(fr, ff, which=w).item())
        out["rf_noreal_" + w] = np.array(ref.loss.real_fake_loss(None, ff, which=w).item())
    p_ = torch.tensor([[0.2], [0.9], [0.5], [0.4]]); t_ = torch.tensor([[0.5], [0.3], [0.5], [0.8]]); e_ = torch.tensor([[1.], [0.], [0.], [1.]])
    out["recon_l1"] = np.array(ref.loss.recon_loss(p_, t_, e_, alpha=0.0, gamma=0.0, norm="l1").item())
    out["recon_l2"] = np.array(ref.loss.recon_loss(p_, t_, e_, alpha=0.3, gamma=0.1, norm="l2").item())
    # (7) random_mask_square_instance keeps whole 16-row regions
    np.random.seed(3)
    bag = torch.arange(160 * 2, dtype=torch.float32).reshape(160, 2) + 1
    mb = ref.func.random_mask_square_instance(bag, 0.5, scale=4, mask_way="mask_zero")
    out["mask_bag_rows_kept"] = (mb.abs().sum(1) > 0).numpy()
    print(f"[golden] misc: ci={out['ci']:.6f} noise_eq={out['noise_eq']}")
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    full = (1024, 384, 384)
    small = (64, 32, 32)
    case_generator(ref, "g_abmil_eval_full", full, 1600, False, 1)
    case_generator(ref, "g_abmil_train_full", full, 640, True, 2)
    case_generator(ref, "g_abmil_eval_small", small, 208, False, 3)
    case_generator(ref, "g_abmil_train_small", small, 96, True, 4, nonneg=True)
    case_disc(ref, "d_rlip_eval_full", 1024, 128, 1600, False, 5)
    case_disc(ref, "d_rlip_train_full", 1024, 128, 640, True, 6)
    case_disc(ref, "d_rlip_eval_small", 64, 32, 208, False, 7)
    case_disc(ref, "d_bag_train_small", 64, 32, 96, True, 8, iprd="bag")
    case_disc_cat(ref, "d_cat_train_full", 1024, 128, 640, True, 13)
    case_disc_cat(ref, "d_cat_eval_small", 64, 32, 208, False, 14)
    case_cluster(ref, "g_cluster_full", full, 800, 9, False)
    case_cluster(ref, "g_cluster_empty_small", small, 160, 10, True)
    case_step(ref, "step_small", small, 32, [96, 160, 48, 208], 11)
    case_step(ref, "step_full", full, 128, [320, 640, 160], 12, n_steps=1)
    case_misc(ref)
    case_esat(ref, "g_esat_eval_full", 1024, 384, 1600, False, 15, True)
    case_esat(ref, "g_esat_train_full", 1024, 384, 640, True, 16, False)
    case_esat(ref, "g_esat_train_small", 64, 32, 208, True, 17, True)
    case_esat(ref, "g_esat_eval_small", 64, 32, 96, False, 18, False)


if __name__ == "__main__":
    main()
