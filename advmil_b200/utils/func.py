"""Host-side helpers that sit ON the hot path in the reference (utils/func.py)."""
from __future__ import annotations

import itertools

import torch

_seed_counter = itertools.count(1)


def generate_noise(*dims, to_device="cpu", distribution="uniform", generator=None):
    """Same RNG stream as the reference's generate_noise (utils/func.py:154-164): Uniform(0,1)/Normal(0,1).sample
    on the CPU default generator is torch.rand / torch.randn of shape dims, then copied to the device."""
    assert distribution in ["uniform", "gaussian"]
    data = torch.rand(*dims, generator=generator) if distribution == "uniform" else torch.randn(*dims, generator=generator)
    return data.to(to_device, non_blocking=True)


def next_dropout_seed() -> int:
    """Seed for the in-kernel counter-based dropout generator.  Derived from the CUDA seed set by seed_everything
    (utils/func.py:166-175) and a call counter, so it neither syncs the device nor consumes the CPU RNG stream that
    the reference uses for the generator noise."""
    base = torch.cuda.initial_seed() if torch.cuda.is_available() else torch.initial_seed()
    return (base * 0x9E3779B97F4A7C15 + next(_seed_counter) * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def sparse_str(s, sep="-", dtype=int):
    if type(s) != str:
        return [s]
    return [dtype(v) for v in s.split(sep)]
