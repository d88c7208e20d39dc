"""GPU: the module surface under the reference handler's own control flow.  The loop below restates
MyHandler._update_disc / _update_gen (model/model_handler.py:349-498) line by line around the advmil_b200 modules —
per-bag calls, boolean-mask indexing of x, pred.detach(), torch losses, torch.optim.Adam built like create_optimizer —
and must reproduce the fixtures that the same loop produced around the reference's modules (tests/golden/step_*.npz)."""
import numpy as np
import pytest
import torch

from oracle import advmil_oracle as O
from tests.util import assert_close, build_D, build_G, d_masks, esat_masks, g_masks, golden, sub, to_dev_masks

pytestmark = pytest.mark.gpu
ZERO_GRAD = ("pool.fc2.bias", "attention_c.bias")


@pytest.mark.parametrize("name", ["step_small", "step_full"])
def test_handler_style_loop_with_dropin_modules(name):
    g = golden(name)
    C, h, o, d, seed, n_steps = [int(v) for v in g["cfg"][:6]]
    Ns = [int(v) for v in g["cfg"][6:]]
    B = len(Ns)
    netG, netD = build_G((C, h, o)), build_D(C, d)
    netG.load_state_dict(O.synth_state_dict(O.G_SHAPES(C, h, o), seed))
    netD.load_state_dict(O.synth_state_dict(O.D_SHAPES(C, d, (64, 128) if d == 128 else (d // 2, d)), seed + 50))
    # optimisers as the handler builds them (model_handler.py:104-107; optim/optim_factory.py:25-37)
    no_decay = [p for n, p in netG.named_parameters() if p.dim() == 1 or n.endswith(".bias")]
    decay = [p for n, p in netG.named_parameters() if not (p.dim() == 1 or n.endswith(".bias"))]
    optG = torch.optim.Adam([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 5e-4}], lr=8e-5)
    optD = torch.optim.Adam(netD.parameters(), lr=8e-5, betas=(0.9, 0.999), weight_decay=0.0)
    xs = [[O.synth_bag(n, seed + i, C).cuda().unsqueeze(0), torch.Tensor([0]).cuda().unsqueeze(0)] for i, n in enumerate(Ns)]
    ys = [torch.tensor([[float(g["t"][i]), float(g["e"][i])]]).cuda() for i in range(B)]
    visible = [bool(v) for v in g["visible"]]
    for step in range(n_steps):
        rng = np.random.default_rng(seed + 100 * step)
        nzD = [torch.tensor(rng.uniform(size=(1, o // 2)), dtype=torch.float32) for _ in range(B)]
        nzG = [torch.tensor(rng.uniform(size=(1, o // 2)), dtype=torch.float32) for _ in range(B)]
        # ---------------- _update_disc ----------------
        netD.train()
        netG.eval()
        reals, fakes = [], []
        for i in range(B):
            data_x, data_x_ext, data_t, data_ind = xs[i][0], xs[i][1], ys[i][:, [0]], ys[i][:, [1]]
            ind_obs = (data_ind == 1).squeeze(-1)
            if torch.sum(ind_obs) > 0 and visible[i]:
                netD._inject_masks = to_dev_masks(d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i))
                reals.append(netD(data_x[ind_obs, :], data_t[ind_obs, :]).view(-1))
            netG.draw_noise = lambda nb, dev, zero, _n=nzD[i]: [None, _n.to(dev)]
            pred = netG(data_x, data_x_ext)
            pat_mask = torch.logical_or(data_ind == 1, data_ind == 0).squeeze(-1)
            netD._inject_masks = to_dev_masks(d_masks(Ns[i] // 16, d, seed + 1000 * step + 10 * i + 5))
            fakes.append(netD(data_x[pat_mask, :], pred[pat_mask, :].detach()).view(-1))
        optD.zero_grad()
        dis_loss = O.real_fake_loss(torch.cat(reals) if reals else None, torch.cat(fakes), "bce")
        dis_loss.backward()
        optD.step()
        # ---------------- _update_gen ----------------
        netD.eval()
        netG.train()
        preds, ff = [], []
        for i in range(B):
            data_x, data_x_ext = xs[i][0], xs[i][1]
            netG._inject_masks = to_dev_masks(g_masks(Ns[i], h, o, seed + 2000 * step + 10 * i))
            netG.draw_noise = lambda nb, dev, zero, _n=nzG[i]: [None, _n.to(dev)]
            pred = netG(data_x, data_x_ext)
            preds.append(pred)
            ff.append(netD(data_x, pred).view(-1))
        optG.zero_grad()
        gen_loss = O.fake_generator_loss(torch.cat(ff))
        vis = [i for i in range(B) if visible[i]]
        t_reg = O.recon_loss(torch.cat([preds[i] for i in vis]), torch.cat([ys[i][:, [0]] for i in vis]),
                             torch.cat([ys[i][:, [1]] for i in vis]), 0.0, 0.0, "l1")
        total = t_reg + 0.004 * gen_loss
        total = total + O.loss_reg_l1(list(netG.parameters()), 1e-5)
        total.backward()          # also flows into D's parameters, like the reference (quirk: wasted work)
        optG.step()
        assert abs(float(dis_loss) - float(g[f"dis_loss{step}"])) < 2e-5
        assert abs(float(gen_loss) - float(g[f"gen_loss{step}"])) < 2e-5
        assert abs(float(t_reg) - float(g[f"t_reg{step}"])) < 2e-5
        assert abs(float(total) - float(g[f"total{step}"])) < 2e-5
        assert_close(torch.cat(preds).detach().cpu().reshape(-1), g[f"pred_g{step}"], 1e-5, f"pred_g step {step}")
        assert_close(torch.cat(ff).detach().cpu(), g[f"fake_g{step}"], 1e-5, f"fake_g step {step}", atol_scale=1e-1)
    for k, p in netG.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(sub(p), g["gparam." + k], 1e-5, "G param " + k, atol=8e-5 * n_steps * 2e-2)
    for k, p in netD.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(sub(p), g["dparam." + k], 1e-5, "D param " + k, atol=8e-5 * n_steps * 2e-2)


def test_handler_style_loop_with_esat_generator():
    """The same control flow with bcb_mode 'patch' (DualTrans_HS generator, RLIP discriminator): two optimiser steps of
    _update_disc / _update_gen around the drop-in modules with torch.optim.Adam, against the oracle's restatement of the
    step over the ESAT generator (losses, per-bag outputs, parameters after the updates).  x_ext is None: the handler's
    placeholder Tensor([0]) makes the reference's own compute_pe raise (model/backbone_utils.py:90-99)."""
    C, d = 1024, 384
    Ns = [160, 320, 96]
    B = len(Ns)
    sdG, sdD = O.synth_state_dict(O.G_ESAT_SHAPES(C, d), 81), O.synth_state_dict(O.D_SHAPES(), 82)
    tr = O.CpuTrainer(sdG, sdD, backbone="patch")
    netG, netD = build_G((C, d, d), mode="patch"), build_D()
    netG.load_state_dict(sdG)
    netD.load_state_dict(sdD)
    no_decay = [p for n, p in netG.named_parameters() if p.dim() == 1 or n.endswith(".bias")]
    decay = [p for n, p in netG.named_parameters() if not (p.dim() == 1 or n.endswith(".bias"))]
    optG = torch.optim.Adam([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 5e-4}], lr=8e-5)
    optD = torch.optim.Adam(netD.parameters(), lr=8e-5, betas=(0.9, 0.999), weight_decay=0.0)
    xs = [O.synth_bag(n, 90 + i, C) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(B, 83)
    es[0] = 1.0
    vis = [True, True, False]
    for step in range(2):
        rng = np.random.default_rng(84 + step)
        nd = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        ng = torch.tensor(rng.uniform(size=(B, d // 2)), dtype=torch.float32)
        mr = [d_masks(n // 16, 128, 300 + 10 * i + step) for i, n in enumerate(Ns)]
        mf = [d_masks(n // 16, 128, 400 + 10 * i + step) for i, n in enumerate(Ns)]
        mg = [esat_masks(n // 16, d, 500 + 10 * i + step) for i, n in enumerate(Ns)]
        ref = tr.step(xs, ts, es, vis, list(nd), list(ng), mr, mf, mg)
        # ---------------- _update_disc ----------------
        netD.train()
        netG.eval()
        reals, fakes = [], []
        for i in range(B):
            data_x, data_t = xs[i].cuda().unsqueeze(0), ts[i].reshape(1, 1).cuda()
            if float(es[i]) == 1 and vis[i]:
                netD._inject_masks = to_dev_masks(mr[i])
                reals.append(netD(data_x, data_t).view(-1))
            netG.draw_noise = lambda nb, dev, zero, _n=nd[i:i + 1]: [None, _n.to(dev)]
            pred = netG(data_x, None)
            netD._inject_masks = to_dev_masks(mf[i])
            fakes.append(netD(data_x, pred.detach()).view(-1))
        optD.zero_grad()
        dis_loss = O.real_fake_loss(torch.cat(reals) if reals else None, torch.cat(fakes), "bce")
        dis_loss.backward()
        optD.step()
        # ---------------- _update_gen ----------------
        netD.eval()
        netG.train()
        preds, ff = [], []
        for i in range(B):
            data_x = xs[i].cuda().unsqueeze(0)
            m = {k: v.to(torch.uint8).contiguous().cuda() for k, v in mg[i].items() if k != "attn"}
            m["attn"] = [mg[i]["attn"].to(torch.uint8).cuda()]
            netG._inject_masks = m
            netG.draw_noise = lambda nb, dev, zero, _n=ng[i:i + 1]: [None, _n.to(dev)]
            pred = netG(data_x, None)
            preds.append(pred)
            ff.append(netD(data_x, pred).view(-1))
        optG.zero_grad()
        gen_loss = O.fake_generator_loss(torch.cat(ff))
        v = [i for i in range(B) if vis[i]]
        t_reg = O.recon_loss(torch.cat([preds[i] for i in v]), ts[v].cuda(), es[v].cuda(), 0.0, 0.0, "l1")
        total = t_reg + 0.004 * gen_loss + O.loss_reg_l1(list(netG.parameters()), 1e-5)
        total.backward()
        optG.step()
        assert abs(float(dis_loss) - ref["dis_loss"]) < 2e-5 and abs(float(gen_loss) - ref["gen_loss"]) < 2e-5
        assert abs(float(t_reg) - ref["t_reg"]) < 2e-5 and abs(float(total) - ref["total"]) < 2e-5
        assert_close(torch.cat(preds).detach().cpu().reshape(-1), ref["pred_g"].reshape(-1), 1e-5, f"pred_g step {step}")
        assert_close(torch.cat(ff).detach().cpu(), ref["fake_g"].reshape(-1), 1e-5, f"fake_g step {step}", atol_scale=1e-1)
    for k, p in netG.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdG[k].detach(), 1e-5, "G param " + k, atol=8e-5 * 2 * 2e-2)
    for k, p in netD.named_parameters():
        if not k.endswith(ZERO_GRAD):
            assert_close(p.detach().cpu(), tr.sdD[k].detach(), 1e-5, "D param " + k, atol=8e-5 * 2 * 2e-2)
