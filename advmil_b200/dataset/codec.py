"""Lossless 12-bit transport format ("p12") of bf16 feature matrices: host-side encoder and the device decoder call.

A bf16 word is sign(1) | exponent(8) | mantissa(7).  The exponents of a feature matrix cluster on a few values, so a
packed step is stored as `lo` (sign << 7 | mantissa, one byte per element), `hi` (two 4-bit exponent codes per byte), a
16-entry code table and a sparse list of escapes for exponents outside the table: 12 bits per element instead of 16,
exact for every bit pattern.  The end-to-end path is bound by the PCIe link, so the 25 % fewer bytes are 25 % less time
per step; decoding costs ~0.2 ms per 16-bag step on the copy stream (csrc/codec.cu)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class P12:
    lo: torch.Tensor        # [n] uint8
    hi: torch.Tensor        # [n/2] uint8
    table: bytes            # 16 bytes (code -> exponent byte; entry 15 unused)
    esc_idx: torch.Tensor   # [m] int32
    esc_exp: torch.Tensor   # [m] uint8
    shape: tuple

    @property
    def nbytes(self) -> int:
        return self.lo.numel() + self.hi.numel() + 16 + 5 * self.esc_idx.numel()

    def pin(self) -> "P12":
        def _p(t):
            try:
                return t.pin_memory()
            except RuntimeError:
                return t
        return P12(_p(self.lo), _p(self.hi), self.table, _p(self.esc_idx), _p(self.esc_exp), self.shape)


def encode_bf16_p12(x: torch.Tensor) -> P12:
    """x: bfloat16 CPU tensor with numel % 8 == 0 (and < 2^31 elements) -> P12 (pageable; call .pin() for the feeder)."""
    assert x.dtype == torch.bfloat16 and not x.is_cuda and x.numel() % 8 == 0 and x.numel() < 2 ** 31
    v = x.contiguous().view(torch.int16).numpy().view(np.uint16).reshape(-1)
    exp = ((v >> 7) & 0xFF).astype(np.uint8)
    hist = np.bincount(exp, minlength=256)
    top = np.argsort(-hist, kind="stable")[:15].astype(np.uint8)
    lut = np.full(256, 15, dtype=np.uint8)
    lut[top] = np.arange(15, dtype=np.uint8)
    codes = lut[exp]
    esc = np.nonzero(codes == 15)[0]
    # an exponent that is in the table never escapes; exponents of rank >= 15 always do
    hi = (codes[0::2] | (codes[1::2] << 4)).astype(np.uint8)
    lo = (((v >> 8) & 0x80) | (v & 0x7F)).astype(np.uint8)
    table = bytes(top.tolist()) + bytes(16 - len(top))
    return P12(torch.from_numpy(lo), torch.from_numpy(hi), table, torch.from_numpy(esc.astype(np.int32)),
               torch.from_numpy(exp[esc].copy()), tuple(x.shape))


def decode_p12_host(p: P12) -> torch.Tensor:
    """Host decoder (numpy) -- the definition the device kernel is tested against."""
    lo, hi = p.lo.numpy().astype(np.uint16), p.hi.numpy()
    codes = np.empty(lo.size, dtype=np.uint8)
    codes[0::2], codes[1::2] = hi & 0xF, hi >> 4
    tab = np.frombuffer(p.table, dtype=np.uint8).astype(np.uint16).copy()
    tab[15] = 0
    exp = tab[codes]
    exp[p.esc_idx.numpy().astype(np.int64)] = p.esc_exp.numpy().astype(np.uint16)
    v = ((lo & 0x80) << 8) | (exp << 7) | (lo & 0x7F)
    return torch.from_numpy(v.astype(np.uint16).view(np.int16)).view(torch.bfloat16).reshape(p.shape)


def decode_p12_device(lo: torch.Tensor, hi: torch.Tensor, table: bytes, esc_idx: torch.Tensor, esc_exp: torch.Tensor,
                      out: torch.Tensor) -> torch.Tensor:
    """All tensors on the device; decodes on the current stream into `out` (bfloat16, lo.numel() elements)."""
    from .. import _lib
    lib = _lib.load()
    assert lo.is_cuda and out.is_cuda and out.dtype == torch.bfloat16 and out.numel() == lo.numel() and out.is_contiguous()
    n_esc = int(esc_idx.numel())
    _lib.check(lib.advmil_bf16p12_decode(lo.data_ptr(), hi.data_ptr(), table, esc_idx.data_ptr() if n_esc else None,
                                         esc_exp.data_ptr() if n_esc else None, lo.numel(), n_esc, out.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "advmil_bf16p12_decode")
    return out


# =====================================================================================================
# "vl": the same transport format with the exponent plane entropy-coded (csrc/codec.cu)
# =====================================================================================================
@dataclass
class VL:
    lo: torch.Tensor        # [n] uint8
    stream: torch.Tensor    # [words] int32 (uint32 bit patterns), incl. 2 guard words
    sbase: torch.Tensor     # [n / 4096] int32
    loff: torch.Tensor      # [n / 128] int16 (uint16 bit patterns)
    tab_exp: np.ndarray     # [16] uint8, host
    tab_len: np.ndarray     # [16] uint8, host
    tab_code: np.ndarray    # [16] uint16, host
    esc_idx: torch.Tensor   # [m] int32
    esc_exp: torch.Tensor   # [m] uint8
    shape: tuple
    blob: Optional[torch.Tensor] = None      # set by pin(): one pinned buffer holding every plane
    blob_offsets: Optional[list] = None

    @property
    def nbytes(self) -> int:
        return self.lo.numel() + 4 * self.stream.numel() + 4 * self.sbase.numel() + 2 * self.loff.numel() + 64 + 5 * self.esc_idx.numel()

    @property
    def bits_per_element(self) -> float:
        return 8.0 * self.nbytes / max(self.lo.numel(), 1)

    def pin(self) -> "VL":
        """One pinned blob [lo | stream | sbase | loff | esc_idx | esc_exp] (256-byte aligned parts): the feeder moves a step's
        features with a single copy; the member tensors become views into the blob."""
        parts = [self.lo, self.stream, self.sbase, self.loff, self.esc_idx, self.esc_exp]
        offs, total = [], 0
        for t in parts:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        try:
            blob = torch.empty(max(total, 256), dtype=torch.uint8).pin_memory()
        except RuntimeError:
            blob = torch.empty(max(total, 256), dtype=torch.uint8)
        views = []
        for t, o in zip(parts, offs):
            nb = t.numel() * t.element_size()
            v = blob[o:o + nb].view(t.dtype)
            v.copy_(t.reshape(-1))
            views.append(v)
        out = VL(views[0], views[1], views[2], views[3], self.tab_exp, self.tab_len, self.tab_code, views[4], views[5], self.shape)
        out.blob, out.blob_offsets = blob, offs
        return out


def _np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def vl_tables(hist256: np.ndarray):
    """(tab_exp, tab_len, tab_code) from an exponent histogram (256 counts), e.g. of a whole packed file."""
    from .. import _lib
    lib = _lib.load()
    h = np.ascontiguousarray(hist256, dtype=np.uint64)
    assert h.size == 256
    te, tl, tc = np.zeros(16, np.uint8), np.zeros(16, np.uint8), np.zeros(16, np.uint16)
    _lib.check(lib.advmil_bf16vl_tables(_np_ptr(h), _np_ptr(te), _np_ptr(tl), _np_ptr(tc)), "advmil_bf16vl_tables")
    return te, tl, tc


def encode_bf16_vl(x: torch.Tensor, tables=None) -> VL:
    """x: bfloat16 CPU tensor with numel % 4096 == 0 (and < 2^31 elements) -> VL (pageable; call .pin() for the feeder).
    The encoder is the library's host function advmil_bf16vl_encode (one sequential pass, ~1 ns per element).
    tables: (tab_exp, tab_len, tab_code) from `vl_tables` to encode with a shared (per-file) table instead of x's own."""
    import ctypes as C
    from .. import _lib
    lib = _lib.load()
    assert x.dtype == torch.bfloat16 and not x.is_cuda and x.numel() % 4096 == 0 and x.numel() < 2 ** 31
    v = x.contiguous().view(torch.int16).numpy().view(np.uint16).reshape(-1)
    n = v.size
    cap = n // 4 + n // 128 + 2
    lo = np.empty(n, dtype=np.uint8)
    stream = np.empty(cap, dtype=np.uint32)
    sbase = np.empty(max(n // 4096, 1), dtype=np.uint32)
    loff = np.empty(max(n // 128, 1), dtype=np.uint16)
    if tables is None:
        te, tl, tc = np.zeros(16, np.uint8), np.zeros(16, np.uint8), np.zeros(16, np.uint16)
        fn = lib.advmil_bf16vl_encode
    else:
        te, tl, tc = (np.ascontiguousarray(t) for t in tables)
        fn = lib.advmil_bf16vl_encode_with_tables
    esc_cap = n
    ei, ee = np.empty(esc_cap, dtype=np.int32), np.empty(esc_cap, dtype=np.uint8)
    words, nesc = C.c_int64(0), C.c_int64(0)
    _lib.check(fn(_np_ptr(v), n, _np_ptr(lo), _np_ptr(stream), cap, _np_ptr(sbase), _np_ptr(loff), _np_ptr(te), _np_ptr(tl), _np_ptr(tc),
                  _np_ptr(ei), _np_ptr(ee), esc_cap, C.byref(words), C.byref(nesc)), "advmil_bf16vl_encode")
    w, m = int(words.value), int(nesc.value)
    return VL(torch.from_numpy(lo), torch.from_numpy(stream[:w].copy().view(np.int32)), torch.from_numpy(sbase[:n // 4096].view(np.int32)),
              torch.from_numpy(loff[:n // 128].view(np.int16)), te, tl, tc, torch.from_numpy(ei[:m].copy()), torch.from_numpy(ee[:m].copy()),
              tuple(x.shape))


def decode_vl_host(p: VL) -> torch.Tensor:
    """Host decoder (the library's sequential reference, advmil_bf16vl_decode_host) -- what the device kernel is tested against."""
    from .. import _lib
    lib = _lib.load()
    n = p.lo.numel()
    out = np.empty(n, dtype=np.uint16)
    m = int(p.esc_idx.numel())
    _lib.check(lib.advmil_bf16vl_decode_host(p.lo.data_ptr(), p.stream.data_ptr(), p.sbase.data_ptr(), p.loff.data_ptr(), _np_ptr(p.tab_exp),
                                             _np_ptr(p.tab_len), _np_ptr(p.tab_code), p.esc_idx.data_ptr() if m else None,
                                             p.esc_exp.data_ptr() if m else None, n, m, _np_ptr(out)), "advmil_bf16vl_decode_host")
    return torch.from_numpy(out.view(np.int16)).view(torch.bfloat16).reshape(p.shape)


def decode_vl_device(lo, stream, sbase, loff, p: VL, esc_idx, esc_exp, out: torch.Tensor) -> torch.Tensor:
    """Planes on the device (tables from `p`, host); decodes on the current stream into `out` (bfloat16, lo.numel() elements)."""
    from .. import _lib
    lib = _lib.load()
    assert lo.is_cuda and out.is_cuda and out.dtype == torch.bfloat16 and out.numel() == lo.numel() and out.is_contiguous()
    m = int(esc_idx.numel())
    _lib.check(lib.advmil_bf16vl_decode(lo.data_ptr(), stream.data_ptr(), sbase.data_ptr(), loff.data_ptr(), _np_ptr(p.tab_exp),
                                        _np_ptr(p.tab_len), _np_ptr(p.tab_code), esc_idx.data_ptr() if m else None,
                                        esc_exp.data_ptr() if m else None, lo.numel(), m, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "advmil_bf16vl_decode")
    return out
