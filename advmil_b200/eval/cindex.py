"""Concordance index on the device, with the reference's entry point and semantics (eval/cindex.py:10-40,106-143):
`concordance_index(y_true, y_pred)` with y_true = [time, event] columns; a scalar prediction per sample is a survival
time (risk = -pred), a row of hazards is a discrete model (risk = -sum_k prod_{l<=k} (1 - h_l)).  The pair counting runs
in libadvmil_b200.so (advmil_cindex_counts, integer exact); the reference's errors are reproduced as ValueError."""
from __future__ import annotations

from typing import Dict, Union

import numpy as np
import torch

from .. import _lib
from .._lib import check


def _dev(v, device, dtype) -> torch.Tensor:
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v))
    return v.to(device=device, dtype=dtype).contiguous()


def _count_dtype(*vs) -> torch.dtype:
    """The reference counts pairs in the dtype its numpy arrays arrive in (eval/cindex.py:106-143): float64 inputs are
    compared as float64, everything else (the handler's float32 tensors) as float32."""
    for v in vs:
        if (isinstance(v, np.ndarray) and v.dtype == np.float64) or (isinstance(v, torch.Tensor) and v.dtype == torch.float64):
            return torch.float64
    return torch.float32


def concordance_counts(t, e, pred, tied_tol: float = 1e-8, device="cuda") -> Dict[str, int]:
    """Integer pair counts: concordant, tied_risk, comparable, discordant (one host sync to read them)."""
    lib = _lib.load()
    dt = _count_dtype(t, pred)
    t, e, pred = _dev(t, device, dt).reshape(-1), _dev(e, device, dt).reshape(-1), _dev(pred, device, dt).reshape(-1)
    n = t.numel()
    if not (e.numel() == n and pred.numel() == n):
        raise ValueError("Found input variables with inconsistent numbers of samples")
    if n < 2:
        raise ValueError("Need a minimum of two samples")
    counts = torch.empty(4, dtype=torch.int64, device=t.device)
    fn = lib.advmil_cindex_counts_f64 if dt == torch.float64 else lib.advmil_cindex_counts
    check(fn(t.data_ptr(), e.data_ptr(), pred.data_ptr(), n, float(tied_tol), counts.data_ptr(),
             torch.cuda.current_stream().cuda_stream), "advmil_cindex_counts")
    c = counts.tolist()
    return {"concordant": c[0], "tied_risk": c[1], "comparable": c[2], "discordant": c[3]}


def concordance_index(y_true: Union[torch.Tensor, np.ndarray], y_pred: Union[torch.Tensor, np.ndarray], device="cuda") -> float:
    if isinstance(y_true, np.ndarray):
        y_true = torch.from_numpy(y_true)
    if isinstance(y_pred, np.ndarray):
        y_pred = torch.from_numpy(y_pred)
    y_true = y_true.reshape(y_true.shape[0], -1) if y_true.dim() != 2 else y_true
    t, e = y_true[:, 0], y_true[:, 1]
    if not bool((e != 0).any()):
        raise ValueError("All samples are censored")
    if y_pred.dim() == 1 or y_pred.shape[1] == 1:
        pred = y_pred.reshape(-1)
    else:   # discrete model: a larger expected survival is a smaller risk, exactly like a larger predicted time
        yp = y_pred.to(device=device, dtype=_count_dtype(y_pred))
        pred = torch.cumprod(1.0 - yp, dim=1).sum(dim=1)
    c = concordance_counts(t, e, pred, 1e-8, device)
    if c["comparable"] == 0:
        raise ValueError("Data has no comparable pairs, cannot estimate concordance index.")
    return (c["concordant"] + 0.5 * c["tied_risk"]) / c["comparable"]
