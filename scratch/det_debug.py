import os, sys, torch
sys.path.insert(0, os.getcwd())
from tests.test_gpu_fullsize import _inputs, _run_step, N_FULL, B_FULL
from oracle import advmil_oracle as O
sdG, sdD = O.synth_state_dict(O.G_SHAPES(), 71), O.synth_state_dict(O.D_SHAPES(), 72)
for nb, n in ((B_FULL, N_FULL), (4, 2048)):
    xs, ts, es, vis, nd, ng = _inputs([n] * nb, 73, torch.bfloat16)
    runs = [_run_step("bf16", sdG, sdD, xs, ts, es, vis, nd, ng, [5, 6]) for _ in range(3)]
    for r in runs[1:]:
        for k in ("pred_d", "pred_g", "f_d", "f_fake_g", "losses"):
            d = (runs[0][0][k] - r[0][k]).abs().max().item()
            print(nb, n, k, "equal" if torch.equal(runs[0][0][k], r[0][k]) else f"DIFF {d:.3e}", runs[0][0][k][:8].tolist() if k == "losses" else "")
        for name, u, v in zip(("Dgrad", "Ggrad", "Gflat", "Dflat"), runs[0][1:], r[1:]):
            print(nb, n, name, "equal" if torch.equal(u, v) else f"DIFF {(u - v).abs().max().item():.3e} of {u.abs().max().item():.3e}")
