"""On-disk packed bag format: ONE mmap-able blob per dataset split instead of one `torch.load` per slide.

The reference reads every slide of a patient with `torch.load`, concatenates them and moves the result to the GPU through
pageable memory, once per bag per epoch (dataset/PatchWSI.py:65-94, utils/io.py:78-101, model/model_handler.py:315-316).
Here the features of all bags live in one file, row-major `[rows, C]` in fp32 or bf16 (the bf16 mode's storage format:
rounded to nearest-even once, at packing time), preceded by int64 bag offsets and the (t, e) labels.  A step's bags are
copied from the page cache straight into a pinned `PinnedStep` buffer (`PackedFile.step`), which `DeviceFeeder` then moves
with one asynchronous H2D copy.

Layout (little endian):
    0    8   magic  b"ADVMILPK"
    8    4   version (1)          12  4  dtype (0 = float32, 1 = bfloat16, 2 = bfloat16 in the 12-bit transport form)
    16   4   C                    20  4  n_bags
    24   8   rows                 32  8  offsets_pos   40  8  labels_pos   48  8  names_pos   56  8  feats_pos
    offsets_pos : int64[n_bags + 1]       labels_pos : float32[n_bags, 2] = (t, e)
    names_pos   : utf-8, one patient id per line (may be empty)
    feats_pos   : 4096-byte aligned, rows * C elements
    dtype 2 (dataset/codec.py; written by `write_packed(..., dtype=torch.bfloat16, transport="p12")`): at feats_pos the `lo`
    plane uint8[rows * C], then (4096-aligned) the `hi` plane uint8[rows * C / 2], then (4096-aligned) the code table
    uint8[16], n_esc int64, the sorted global escape indices int64[n_esc] and their exponent bytes uint8[n_esc].  One table
    for the whole file, so the planes of any selection of bags concatenate into a step without re-encoding.
    dtype 3 (`transport="vl"`, the entropy-coded form of dataset/codec.py, ~10.9 bits per element): at feats_pos the `lo` plane
    uint8[rows * C]; then (4096-aligned each) one page holding int64 tail_pos, the Huffman stream uint32[words] (bag after bag, every bag encoded on its own
    with the FILE's tables and closed by two guard words), sbase uint32[rows * C / 4096] (word offset of a super-block inside
    ITS BAG's stream), loff uint16[rows * C / 128]; then the tail: int64 words, int64 bag_words[n_bags + 1] (word offset of
    every bag's stream), tab_exp uint8[16], tab_len uint8[16], tab_code uint16[16], n_esc int64, escape indices int64[n_esc]
    (global element index), exponent bytes uint8[n_esc].  rows * C of every bag must be a multiple of 4096 (C = 1024 and 16-row
    regions: always).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .codec import P12, VL
from .packed import PinnedStep, _pin

MAGIC = b"ADVMILPK"
VERSION = 1
_HDR = struct.Struct("<8sIIIIQQQQQ")
_DT = {torch.float32: 0, torch.bfloat16: 1}
_NP = {0: np.float32, 1: np.uint16}
_TORCH = {0: torch.float32, 1: torch.bfloat16}
PAGE = 4096


def write_packed(path: str, bags: Iterable[torch.Tensor], labels: Sequence[Sequence[float]],
                 dtype: torch.dtype = torch.float32, names: Optional[Sequence[str]] = None,
                 require_multiple_of: int = 16, transport: str = "raw") -> Dict[str, int]:
    """Streams `bags` ([N_i, C] tensors, e.g. the torch.cat of a patient's slides, PatchWSI.py:79) into `path`.
    transport="p12" (bf16 only) stores the lossless 12-bit form that the feeder copies to the device as is."""
    assert dtype in _DT, "features are stored as float32 or bfloat16"
    assert transport in ("raw", "p12", "vl")
    if transport in ("p12", "vl"):
        assert dtype == torch.bfloat16, "the transport forms pack bf16 features"
        info = write_packed(path + ".raw", bags, labels, dtype, names, require_multiple_of)
        try:
            return (_convert_to_p12 if transport == "p12" else _convert_to_vl)(path + ".raw", path, info)
        finally:
            os.remove(path + ".raw")
    labels = np.asarray(labels, dtype=np.float32).reshape(-1, 2)
    n_bags = labels.shape[0]
    names_blob = ("\n".join(names)).encode() if names is not None else b""
    if names is not None:
        assert len(names) == n_bags
    offsets_pos = _HDR.size
    labels_pos = offsets_pos + 8 * (n_bags + 1)
    names_pos = labels_pos + labels.nbytes
    feats_pos = (names_pos + len(names_blob) + PAGE - 1) // PAGE * PAGE
    offsets = [0]
    C = None
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.seek(feats_pos)
        count = 0
        for b in bags:
            b = b[0] if b.dim() == 3 else b
            assert b.dim() == 2, "a bag is [N, C]"
            if C is None:
                C = int(b.shape[1])
            assert b.shape[1] == C, "all bags share the feature width"
            n = int(b.shape[0])
            assert n > 0, f"bag {count} is empty"
            if require_multiple_of:
                assert n % require_multiple_of == 0, \
                    f"bag {count} has {n} instances: the RLIP discriminator needs a multiple of 16 (model/backbone_utils.py:65)"
            v = b.detach().to("cpu", torch.float32).contiguous()
            if dtype == torch.bfloat16:
                f.write(v.to(torch.bfloat16).view(torch.int16).numpy().tobytes())
            else:
                f.write(v.numpy().tobytes())
            offsets.append(offsets[-1] + n)
            count += 1
        assert count == n_bags, f"{count} bags for {n_bags} labels"
        f.seek(0)
        f.write(_HDR.pack(MAGIC, VERSION, _DT[dtype], C, n_bags, offsets[-1], offsets_pos, labels_pos, names_pos, feats_pos))
        f.write(np.asarray(offsets, dtype=np.int64).tobytes())
        f.write(labels.tobytes())
        f.write(names_blob)
    os.replace(tmp, path)
    return {"bags": n_bags, "rows": offsets[-1], "C": C, "bytes": os.path.getsize(path)}


def _convert_to_p12(raw_path: str, path: str, info: Dict[str, int], chunk_rows: int = 1 << 15) -> Dict[str, int]:
    """Second pass of write_packed(transport="p12"): one exponent histogram over the whole file -> one code table."""
    src = PackedFile(raw_path)
    C, rows = src.C, src.rows
    assert (rows * C) % 8 == 0
    hist = np.zeros(256, dtype=np.int64)
    for a in range(0, rows, chunk_rows):
        v = np.asarray(src._mm[a:a + chunk_rows]).reshape(-1)
        hist += np.bincount(((v >> 7) & 0xFF).astype(np.uint8), minlength=256)
    top = np.argsort(-hist, kind="stable")[:15].astype(np.uint8)
    lut = np.full(256, 15, dtype=np.uint8)
    lut[top] = np.arange(15, dtype=np.uint8)
    with open(raw_path, "rb") as f:
        head = f.read(src._feats_pos)
    lo_pos = src._feats_pos
    hi_pos = (lo_pos + rows * C + PAGE - 1) // PAGE * PAGE
    tail_pos = (hi_pos + rows * C // 2 + PAGE - 1) // PAGE * PAGE
    esc_idx, esc_exp = [], []
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(head)
        for a in range(0, rows, chunk_rows):
            v = np.asarray(src._mm[a:a + chunk_rows]).reshape(-1)
            exp = ((v >> 7) & 0xFF).astype(np.uint8)
            codes = lut[exp]
            e = np.nonzero(codes == 15)[0]
            esc_idx.append(e.astype(np.int64) + a * C)
            esc_exp.append(exp[e])
            f.seek(lo_pos + a * C)
            f.write((((v >> 8) & 0x80) | (v & 0x7F)).astype(np.uint8).tobytes())
            f.seek(hi_pos + a * C // 2)
            f.write((codes[0::2] | (codes[1::2] << 4)).astype(np.uint8).tobytes())
        ei, ee = np.concatenate(esc_idx), np.concatenate(esc_exp)
        f.seek(tail_pos)
        f.write(bytes(top.tolist()) + bytes(16 - len(top)))
        f.write(struct.pack("<q", ei.size))
        f.write(ei.tobytes())
        f.write(ee.tobytes())
        f.seek(12)
        f.write(struct.pack("<I", 2))
    os.replace(tmp, path)
    return {**info, "bytes": os.path.getsize(path), "escapes": int(ei.size)}


def _align(v: int) -> int:
    return (v + PAGE - 1) // PAGE * PAGE


def _convert_to_vl(raw_path: str, path: str, info: Dict[str, int], chunk_rows: int = 1 << 15) -> Dict[str, int]:
    """Second pass of write_packed(transport="vl"): one exponent histogram over the whole file -> one Huffman table; every
    bag is then encoded on its own with that table (its sub-stream offsets are relative to the bag)."""
    from .codec import encode_bf16_vl, vl_tables
    src = PackedFile(raw_path)
    C, rows = src.C, src.rows
    hist = np.zeros(256, dtype=np.int64)
    for a in range(0, rows, chunk_rows):
        v = np.asarray(src._mm[a:a + chunk_rows]).reshape(-1)
        hist += np.bincount(((v >> 7) & 0xFF).astype(np.uint8), minlength=256)
    tabs = vl_tables(hist)
    with open(raw_path, "rb") as f:
        head = f.read(src._feats_pos)
    n = rows * C
    lo_pos = src._feats_pos
    dir_pos = _align(lo_pos + n)            # one page: int64 tail_pos
    stream_pos = dir_pos + PAGE
    tmp = path + ".tmp"
    sbase_all, loff_all, bag_words, esc_idx, esc_exp = [], [], [0], [], []
    with open(tmp, "wb") as f:
        f.write(head)
        f.seek(stream_pos)
        for i in range(src.n_bags):
            a, b = int(src.offsets[i]), int(src.offsets[i + 1])
            assert ((b - a) * C) % 4096 == 0, f"bag {i}: rows * C must be a multiple of 4096 for the vl form"
            x = torch.from_numpy(np.array(src._mm[a:b])).view(torch.bfloat16)
            p = encode_bf16_vl(x, tables=tabs)
            f.write(p.stream.numpy().tobytes())
            bag_words.append(bag_words[-1] + p.stream.numel())
            sbase_all.append(p.sbase.numpy().view(np.uint32))
            loff_all.append(p.loff.numpy().view(np.uint16))
            esc_idx.append(p.esc_idx.numpy().astype(np.int64) + a * C)
            esc_exp.append(p.esc_exp.numpy())
            pos = f.tell()
            f.seek(lo_pos + a * C)
            f.write(p.lo.numpy().tobytes())
            f.seek(pos)
        words = bag_words[-1]
        sbase_pos = _align(stream_pos + 4 * words)
        loff_pos = _align(sbase_pos + 4 * (n // 4096))
        tail_pos = _align(loff_pos + 2 * (n // 128))
        f.seek(sbase_pos)
        f.write(np.concatenate(sbase_all).astype(np.uint32).tobytes())
        f.seek(loff_pos)
        f.write(np.concatenate(loff_all).astype(np.uint16).tobytes())
        ei, ee = np.concatenate(esc_idx), np.concatenate(esc_exp)
        f.seek(tail_pos)
        f.write(struct.pack("<q", words))
        f.write(np.asarray(bag_words, dtype=np.int64).tobytes())
        f.write(tabs[0].tobytes() + tabs[1].tobytes() + tabs[2].astype(np.uint16).tobytes())
        f.write(struct.pack("<q", ei.size))
        f.write(ei.tobytes())
        f.write(ee.astype(np.uint8).tobytes())
        f.seek(dir_pos)
        f.write(struct.pack("<q", tail_pos))
        f.seek(12)
        f.write(struct.pack("<I", 3))
    os.replace(tmp, path)
    return {**info, "bytes": os.path.getsize(path), "escapes": int(ei.size), "bits_per_element": 8.0 * (os.path.getsize(path) - lo_pos) / n}


def pack_reference_layout(patient_slides: Dict[str, List[str]], labels: Dict[str, Tuple[float, float]], out_path: str,
                          dtype: torch.dtype = torch.float32, trim_to_multiple_of: int = 16, transport: str = "raw") -> Dict[str, int]:
    """Converts the reference's per-slide `.pt` feature files (utils/io.py:78-101 `read_patch_data`, one [n, C] tensor per
    slide) into one packed file: per patient the slides are concatenated in the given order (PatchWSI.py:74-79).
    trim_to_multiple_of drops the trailing rows that do not fill a 16-row region (level-1 features produced by
    tools/big_to_small_patching.py are already multiples of 16)."""
    pids = list(patient_slides.keys())

    def gen():
        for pid in pids:
            feats = [torch.load(p, map_location="cpu") for p in patient_slides[pid]]
            x = torch.cat([f.to(torch.float32) for f in feats], dim=0)
            if trim_to_multiple_of:
                x = x[: x.shape[0] // trim_to_multiple_of * trim_to_multiple_of]
            yield x

    return write_packed(out_path, gen(), [labels[p] for p in pids], dtype=dtype, names=pids,
                        require_multiple_of=trim_to_multiple_of, transport=transport)


class PackedFile:
    """Read side: zero-copy views of the bags (np.memmap) and pinned per-step buffers for the device feeder."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            hdr = f.read(_HDR.size)
            magic, ver, dt, C, n_bags, rows, opos, lpos, npos, fpos = _HDR.unpack(hdr)
            if magic != MAGIC or ver != VERSION:
                raise ValueError(f"{path}: not an advmil_b200 packed file (magic {magic!r}, version {ver})")
            f.seek(opos)
            self.offsets = np.frombuffer(f.read(8 * (n_bags + 1)), dtype=np.int64).copy()
            f.seek(lpos)
            self.labels = np.frombuffer(f.read(8 * n_bags), dtype=np.float32).reshape(n_bags, 2).copy()
            f.seek(npos)
            blob = f.read(fpos - npos).rstrip(b"\x00")
            self.names = blob.decode().split("\n") if blob else None
        need = rows * C * 4 if dt == 0 else rows * C * 2 if dt == 1 else rows * C * 3 // 2 if dt == 2 else rows * C
        if int(self.offsets[-1]) != rows or os.path.getsize(path) < fpos + need:
            raise ValueError(f"{path}: truncated or inconsistent packed file")
        self.C, self.n_bags, self.rows, self.code = int(C), int(n_bags), int(rows), int(dt)
        self._feats_pos = int(fpos)
        self.dtype = torch.bfloat16 if dt in (2, 3) else _TORCH[dt]
        if dt == 3:     # entropy-coded transport form: lo plane + per-bag Huffman streams + file-wide tables
            n = self.rows * self.C
            dir_pos = _align(fpos + n)
            stream_pos = dir_pos + PAGE
            with open(path, "rb") as f:
                f.seek(dir_pos)
                tail_pos = struct.unpack("<q", f.read(8))[0]
                f.seek(tail_pos)
                words = struct.unpack("<q", f.read(8))[0]
                self._bag_words = np.frombuffer(f.read(8 * (n_bags + 1)), dtype=np.int64).copy()
                self._tabs = (np.frombuffer(f.read(16), dtype=np.uint8).copy(), np.frombuffer(f.read(16), dtype=np.uint8).copy(),
                              np.frombuffer(f.read(32), dtype=np.uint16).copy())
                n_esc = struct.unpack("<q", f.read(8))[0]
                self._esc_idx = np.frombuffer(f.read(8 * n_esc), dtype=np.int64).copy()
                self._esc_exp = np.frombuffer(f.read(n_esc), dtype=np.uint8).copy()
            sbase_pos = _align(stream_pos + 4 * words)
            loff_pos = _align(sbase_pos + 4 * (n // 4096))
            self._lo = np.memmap(path, dtype=np.uint8, mode="r", offset=fpos, shape=(self.rows, self.C))
            self._stream = np.memmap(path, dtype=np.int32, mode="r", offset=stream_pos, shape=(int(words),))
            self._sbase = np.memmap(path, dtype=np.int32, mode="r", offset=sbase_pos, shape=(n // 4096,))
            self._loff = np.memmap(path, dtype=np.int16, mode="r", offset=loff_pos, shape=(n // 128,))
            self._mm = None
        elif dt == 2:     # 12-bit transport form: two byte planes + table + sorted global escapes
            hi_pos = (fpos + rows * C + PAGE - 1) // PAGE * PAGE
            tail_pos = (hi_pos + rows * C // 2 + PAGE - 1) // PAGE * PAGE
            self._lo = np.memmap(path, dtype=np.uint8, mode="r", offset=fpos, shape=(self.rows, self.C))
            self._hi = np.memmap(path, dtype=np.uint8, mode="r", offset=hi_pos, shape=(self.rows, self.C // 2))
            with open(path, "rb") as f:
                f.seek(tail_pos)
                self._table = f.read(16)
                n_esc = struct.unpack("<q", f.read(8))[0]
                self._esc_idx = np.frombuffer(f.read(8 * n_esc), dtype=np.int64).copy()
                self._esc_exp = np.frombuffer(f.read(n_esc), dtype=np.uint8).copy()
            self._mm = None
        else:
            self._mm = np.memmap(path, dtype=_NP[dt], mode="r", offset=fpos, shape=(self.rows, self.C))

    def __len__(self) -> int:
        return self.n_bags

    @property
    def lengths(self) -> List[int]:
        return [int(b - a) for a, b in zip(self.offsets[:-1], self.offsets[1:])]

    def bag(self, i: int) -> torch.Tensor:
        """[N_i, C] copy of bag i in the stored dtype (what WSIPatch.__getitem__ returns as `feats`, PatchWSI.py:79)."""
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        if self.code == 2:
            from .codec import decode_p12_host
            return decode_p12_host(self._p12([i], pin=False))
        if self.code == 3:
            from .codec import decode_vl_host
            return decode_vl_host(self._vl([i]))
        v = torch.from_numpy(np.array(self._mm[a:b]))
        return v.view(torch.bfloat16) if self.code == 1 else v

    def _p12(self, indices: Sequence[int], pin: bool) -> P12:
        """The 12-bit planes of the bags `indices`, concatenated in step order (escape indices re-based to the step)."""
        lens = [int(self.offsets[i + 1] - self.offsets[i]) for i in indices]
        rows = sum(lens)
        lo, hi = torch.empty(rows * self.C, dtype=torch.uint8), torch.empty(rows * self.C // 2, dtype=torch.uint8)
        if pin:
            lo, hi = _pin(lo), _pin(hi)
        lo_np, hi_np = lo.numpy().reshape(rows, self.C), hi.numpy().reshape(rows, self.C // 2)
        ei, ee, off = [], [], 0
        for i, n in zip(indices, lens):
            a = int(self.offsets[i])
            np.copyto(lo_np[off:off + n], self._lo[a:a + n])
            np.copyto(hi_np[off:off + n], self._hi[a:a + n])
            l, r = np.searchsorted(self._esc_idx, [a * self.C, (a + n) * self.C])
            ei.append(self._esc_idx[l:r] - a * self.C + off * self.C)
            ee.append(self._esc_exp[l:r])
            off += n
        ei = torch.from_numpy(np.concatenate(ei).astype(np.int32))
        ee = torch.from_numpy(np.concatenate(ee).copy())
        mk = _pin if pin else (lambda v: v)
        return P12(lo, hi, self._table, mk(ei), mk(ee), (rows, self.C))

    def _vl(self, indices: Sequence[int]) -> VL:
        """The entropy-coded form of the bags `indices`, concatenated in step order: lo planes and streams as stored, the
        super-block offsets re-based to the step's stream, escape indices re-based to the step (pageable; VL.pin() makes the
        feeder's single pinned blob)."""
        lens = [int(self.offsets[i + 1] - self.offsets[i]) for i in indices]
        rows, C = sum(lens), self.C
        lo = np.empty(rows * C, dtype=np.uint8)
        streams, sbases, loffs, ei, ee, off, woff = [], [], [], [], [], 0, 0
        for i, n in zip(indices, lens):
            a = int(self.offsets[i])
            lo[off * C:(off + n) * C] = self._lo[a:a + n].reshape(-1)
            w0, w1 = int(self._bag_words[i]), int(self._bag_words[i + 1])
            streams.append(np.asarray(self._stream[w0:w1]))
            sbases.append(np.asarray(self._sbase[a * C // 4096:(a + n) * C // 4096]).astype(np.int64) + woff)
            loffs.append(np.asarray(self._loff[a * C // 128:(a + n) * C // 128]))
            l, r = np.searchsorted(self._esc_idx, [a * C, (a + n) * C])
            ei.append(self._esc_idx[l:r] - a * C + off * C)
            ee.append(self._esc_exp[l:r])
            off += n
            woff += w1 - w0
        return VL(torch.from_numpy(lo), torch.from_numpy(np.concatenate(streams).astype(np.int32)),
                  torch.from_numpy(np.concatenate(sbases).astype(np.uint32).view(np.int32)), torch.from_numpy(np.concatenate(loffs).astype(np.int16)),
                  self._tabs[0], self._tabs[1], self._tabs[2], torch.from_numpy(np.concatenate(ei).astype(np.int32)),
                  torch.from_numpy(np.concatenate(ee).copy()), (rows, C))

    def step(self, indices: Sequence[int], visible: Optional[Sequence[bool]] = None, pin: bool = True) -> PinnedStep:
        """The bags `indices` of one optimiser step as a pinned, packed buffer (page cache -> pinned memory, one copy)."""
        lens = [int(self.offsets[i + 1] - self.offsets[i]) for i in indices]
        p12 = vl = None
        if self.code == 2:      # the planes go to the device as stored; the bf16 matrix only exists there
            p12 = self._p12(indices, pin)
            x = torch.empty(0, self.C, dtype=torch.bfloat16)
        elif self.code == 3:
            vl = self._vl(indices)
            vl = vl.pin() if pin else vl
            x = torch.empty(0, self.C, dtype=torch.bfloat16)
        else:
            x = torch.empty(sum(lens), self.C, dtype=self.dtype)
            if pin:
                x = _pin(x)
            dst = x.view(torch.int16).numpy() if self.code == 1 else x.numpy()
            off = 0
            for i, n in zip(indices, lens):
                a = int(self.offsets[i])
                np.copyto(dst[off:off + n], self._mm[a:a + n].view(np.int16) if self.code == 1 else self._mm[a:a + n])
                off += n
        lab = torch.from_numpy(self.labels[list(indices)].copy())
        mk = _pin if pin else (lambda v: v)
        vis = [True] * len(lens) if visible is None else list(visible)
        return PinnedStep(x=x, lengths=lens, t=mk(lab[:, 0].contiguous()), e=mk(lab[:, 1].contiguous()),
                          idx=torch.tensor(list(indices), dtype=torch.int32),
                          visible=mk(torch.tensor([1 if v else 0 for v in vis], dtype=torch.uint8)), p12=p12, vl=vl)
