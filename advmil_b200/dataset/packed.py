"""Packed, pinned, offset-indexed bag storage and an asynchronous device feeder.

Rebuilds the role of the reference's WSIPatch dataset + DataLoader + `.cuda()` (dataset/PatchWSI.py:65-94;
model/model_handler.py:158-165,315-316) for the B200 path:

  reference : per-slide torch.load -> torch.cat per patient -> default_collate -> pageable, synchronous H2D per bag
  here      : bags of one optimiser step are packed once into a pinned [rows, C] buffer + int32 offsets
              (`PinnedStep`), and `DeviceFeeder` copies step k+1 on a side stream while step k computes
              (ring of device buffers, event hand-off, no host sync in the steady state).

Patch mode yields (feats, zeros(1)) and cluster mode (feats, cluster ids) like WSIPatch.__getitem__ (:82-94).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from ..ops import PackedBags
from .codec import P12, VL, decode_p12_device, decode_vl_device


@dataclass
class PinnedStep:
    """The bags of one optimiser step (`bp_every_batch` of them, config/cfg_nlst.yaml:71) in pinned host memory."""
    x: torch.Tensor            # [rows, C] float32 (or bfloat16: the bf16 mode's storage format), pinned
    lengths: List[int]
    t: torch.Tensor            # [bags] float32, pinned
    e: torch.Tensor            # [bags] float32, pinned
    idx: torch.Tensor          # [bags] int32 (dataset indices)
    visible: torch.Tensor      # [bags] uint8 (label_visible_mask, model_handler.py:591-596)
    cluster_id: Optional[torch.Tensor] = None   # [rows] int32 (cluster mode)
    offsets: Optional[torch.Tensor] = None      # [bags+1] int32 prefix sums of lengths, pinned
    p12: Optional["P12"] = None                 # lossless 12-bit transport form of x (bf16 only): what the feeder copies
    vl: Optional["VL"] = None                   # the same with the exponent plane entropy-coded (~10.9 bits per element)

    def __post_init__(self):
        if self.offsets is None:
            offs = [0]
            for n in self.lengths:
                offs.append(offs[-1] + int(n))
            self.offsets = _pin(torch.tensor(offs, dtype=torch.int32))

    @property
    def nbytes(self) -> int:
        """Bytes the feeder copies to the device for this step."""
        feat = self.vl.nbytes if self.vl is not None else self.p12.nbytes if self.p12 is not None else self.x.numel() * self.x.element_size()
        return feat + (self.t.numel() + self.e.numel()) * 4 + self.visible.numel()

    def pack12(self) -> "PinnedStep":
        """Adds the 12-bit transport form of the bf16 features (dataset/codec.py); the feeder then copies 12 instead of
        16 bits per element and decodes on the device.  Exact: the device sees the same bf16 words."""
        from .codec import encode_bf16_p12
        assert self.x.dtype == torch.bfloat16, "the 12-bit transport format packs bf16 features"
        if self.p12 is None:
            self.p12 = encode_bf16_p12(self.x).pin()
        return self


def _packvl(self) -> "PinnedStep":
    """Adds the entropy-coded transport form of the bf16 features (dataset/codec.py, "vl"): ~10.9 instead of 12 bits per element
    on Gaussian-like features; takes precedence over p12 in the feeder.  Exact: the device sees the same bf16 words."""
    from .codec import encode_bf16_vl
    assert self.x.dtype == torch.bfloat16, "the transport formats pack bf16 features"
    assert self.x.numel() % 4096 == 0, "the vl transport form needs a multiple of 4096 elements per step"
    if self.vl is None:
        self.vl = encode_bf16_vl(self.x).pin()
    return self


PinnedStep.packvl = _packvl


def bind_host_to_gpu(device_index: int, ranks_on_node: int = 1, local_rank: int = 0) -> dict:
    """Pins the calling process to the CPU cores of the NUMA node its GPU hangs off (NVML's ideal CPU affinity) BEFORE any
    pinned staging buffer is allocated: with the kernel's first-touch policy the pinned arena then lives in the memory
    that is local to the GPU's PCIe root, instead of wherever the launcher happened to start the rank.  Without this,
    eight ranks of one node read their features out of one socket's memory (round 1: 22 GB/s per GPU at N = 8 against 55
    GB/s alone).  When several ranks share a node's core set the set is split evenly between them.  Returns what it did;
    never raises (a box without NVML / sched_setaffinity keeps the inherited affinity)."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip() != ""]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        # ranks whose GPUs share this core set split it (same NVML mask => same set): contiguous slices by local rank
        share = max(1, min(ranks_on_node, len(allowed)))
        per = max(1, len(allowed) // share)
        k = local_rank % share
        mine = allowed[k * per:(k + 1) * per] if share > 1 else allowed
        os.sched_setaffinity(0, mine or allowed)
        info.update(bound=True, cpus=len(mine or allowed), first_cpu=(mine or allowed)[0], numa_cpus=len(allowed))
    except Exception as ex:       # noqa: BLE001
        info["error"] = str(ex)[:120]
    return info


def _pin(t: torch.Tensor) -> torch.Tensor:
    try:
        return t.pin_memory()
    except RuntimeError:      # no CUDA runtime (CPU-only container): keep pageable memory, same layout
        return t


def pack_step(bags: Sequence[torch.Tensor], labels: Sequence[Sequence[float]], idx: Optional[Sequence[int]] = None,
              visible: Optional[Sequence[bool]] = None, cluster_ids: Optional[Sequence[torch.Tensor]] = None,
              require_multiple_of: int = 16, pin: bool = True, dtype: torch.dtype = torch.float32) -> PinnedStep:
    """Packs per-patient feature tensors ([N_i, C], as WSIPatch returns them after torch.cat, :79) into one buffer.
    dtype=torch.bfloat16 stores the features in the bf16 mode's format (rounded once, at packing time: half the host
    memory and half the H2D bytes of every step)."""
    assert dtype in (torch.float32, torch.bfloat16)
    assert len(bags) == len(labels) and len(bags) > 0
    lengths = [int(b.shape[0]) for b in bags]
    for i, n in enumerate(lengths):
        assert n > 0, f"bag {i} is empty"
        if require_multiple_of:
            assert n % require_multiple_of == 0, \
                f"bag {i} has {n} instances: the RLIP discriminator needs a multiple of 16 (model/backbone_utils.py:65)"
    C = int(bags[0].shape[1])
    rows = sum(lengths)
    x = torch.empty(rows, C, dtype=dtype)
    if pin:
        x = _pin(x)
    off = 0
    for b in bags:
        assert b.shape[1] == C
        x[off:off + b.shape[0]].copy_(b.to(torch.float32))     # dataset/PatchWSI.py:79 `.to(torch.float)` (+ rounding to dtype)
        off += b.shape[0]
    lab = torch.tensor(np.asarray(labels, dtype=np.float32).reshape(len(bags), 2))
    mk = (lambda v: _pin(v)) if pin else (lambda v: v)
    cid = None
    if cluster_ids is not None:
        cid = mk(torch.cat([torch.as_tensor(c).reshape(-1).to(torch.int32) for c in cluster_ids]))
        assert cid.shape[0] == rows, "one cluster id per instance (dataset/PatchWSI.py:93)"
    return PinnedStep(
        x=x, lengths=lengths, t=mk(lab[:, 0].contiguous()), e=mk(lab[:, 1].contiguous()),
        idx=torch.tensor(list(range(len(bags))) if idx is None else list(idx), dtype=torch.int32),
        visible=mk(torch.tensor([1 if (visible is None or v) else 0 for v in (visible or [True] * len(bags))],
                                dtype=torch.uint8)),
        cluster_id=cid)


def group_steps(n_items: int, bags_per_step: int) -> List[List[int]]:
    """Reference batching: bags are collected 16 at a time; a trailing partial group is dropped each epoch
    (model_handler.py:321)."""
    return [list(range(s, s + bags_per_step)) for s in range(0, n_items - bags_per_step + 1, bags_per_step)]


def shard_bags_balanced(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Data-parallel partition of one step's bags by greedy row balancing (bag sizes vary 100x).
    Deterministic: largest first, ties by index; every rank gets at least one bag when len(lengths) >= world."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        empties = [r for r in range(world) if not out[r]]
        remaining = len(lengths) - sum(len(o) for o in out)
        if empties and remaining <= len(empties):
            r = empties[0]
        else:
            r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += lengths[i]
    return [sorted(o) for o in out]


@dataclass
class DeviceStep:
    bags: PackedBags
    t: torch.Tensor
    e: torch.Tensor
    visible: torch.Tensor
    idx: torch.Tensor
    cluster_id: Optional[torch.Tensor]
    h2d_bytes: int
    counts: tuple = (0.0, 0.0, 0.0)   # local (n_real, n_fake, n_visible) from the host labels (no device sync needed)
    _slot: int = -1


class DeviceFeeder:
    """Double-buffered asynchronous H2D of PinnedSteps: `for step in DeviceFeeder(steps): ...`.

    Slot s holds a device buffer sized for the largest step; the copy of item k+depth-1 is issued on a side stream as soon
    as the consumer releases slot (k-1) (an event recorded on the compute stream when the next item is requested)."""

    def __init__(self, steps: Iterable[PinnedStep], device="cuda", depth: int = 2):
        self.steps = steps
        self.device = torch.device(device)
        self.depth = max(2, depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.decode_stream = torch.cuda.Stream(device=self.device)    # p12 decode runs beside the next step's copies
        self._copied = [torch.cuda.Event() for _ in range(self.depth)]
        self._xbuf: List[Optional[torch.Tensor]] = [None] * self.depth
        self._pbuf: List[Optional[tuple]] = [None] * self.depth      # device staging of the 12-bit planes (lo, hi)
        self._vbuf: List[Optional[torch.Tensor]] = [None] * self.depth   # device staging of the vl form (one blob)
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free = [torch.cuda.Event() for _ in range(self.depth)]

    def _issue(self, slot: int, st: PinnedStep) -> DeviceStep:
        rows, C = (st.vl.shape if st.vl is not None else st.p12.shape if st.p12 is not None else st.x.shape)     # a p12-file step carries no host bf16 matrix
        buf = self._xbuf[slot]
        if buf is None or buf.shape[0] < rows or buf.shape[1] != C or buf.dtype != st.x.dtype:
            buf = torch.empty(rows, C, dtype=st.x.dtype, device=self.device)
            self._xbuf[slot] = buf
        with torch.cuda.stream(self.copy_stream):
            # labels and offsets first (pinned, asynchronous: nothing here blocks the host), then the features
            t = st.t.to(self.device, non_blocking=True)
            e = st.e.to(self.device, non_blocking=True)
            vis = st.visible.to(self.device, non_blocking=True)
            offs = st.offsets.to(self.device, non_blocking=True)
            cid = None if st.cluster_id is None else st.cluster_id.to(self.device, non_blocking=True)
            self.copy_stream.wait_event(self._free[slot])          # consumer finished with this slot
            xd = buf[:rows]
            if st.vl is not None:      # entropy-coded exponent plane: ~10.9 bits per element over the link, ONE copy per step
                n = rows * C
                v = st.vl
                if v.blob is None:
                    v = st.vl = v.pin()
                db = self._vbuf[slot]
                if db is None or db.numel() < v.blob.numel():
                    db = torch.empty(v.blob.numel() + v.blob.numel() // 16, dtype=torch.uint8, device=self.device)
                    self._vbuf[slot] = db
                db[:v.blob.numel()].copy_(v.blob, non_blocking=True)

                def part(i, t):
                    o = v.blob_offsets[i]
                    return db[o:o + t.numel() * t.element_size()].view(t.dtype)
                lo, sw, sbs, lof, ei, ee = (part(i, t) for i, t in enumerate((v.lo, v.stream, v.sbase, v.loff, v.esc_idx, v.esc_exp)))
                self._copied[slot].record(self.copy_stream)
                with torch.cuda.stream(self.decode_stream):
                    self.decode_stream.wait_event(self._copied[slot])
                    decode_vl_device(lo, sw, sbs, lof, v, ei, ee, xd)
                    self._ready[slot].record(self.decode_stream)
            elif st.p12 is None:
                xd.copy_(st.x, non_blocking=True)
            else:       # 12 bits per element over the link; decoded on a second stream so the next copies are not held up
                n = rows * C
                pb = self._pbuf[slot]
                if pb is None or pb[0].numel() < n:
                    pb = (torch.empty(n, dtype=torch.uint8, device=self.device), torch.empty(n // 2, dtype=torch.uint8, device=self.device))
                    self._pbuf[slot] = pb
                lo, hi = pb[0][:n], pb[1][:n // 2]
                ei = st.p12.esc_idx.to(self.device, non_blocking=True)
                ee = st.p12.esc_exp.to(self.device, non_blocking=True)
                lo.copy_(st.p12.lo, non_blocking=True)
                hi.copy_(st.p12.hi, non_blocking=True)
                self._copied[slot].record(self.copy_stream)
                with torch.cuda.stream(self.decode_stream):
                    self.decode_stream.wait_event(self._copied[slot])
                    decode_p12_device(lo, hi, st.p12.table, ei, ee, xd)
                    for ten in (ei, ee):
                        ten.record_stream(self.decode_stream)
                    self._ready[slot].record(self.decode_stream)
            bags = PackedBags(xd, st.lengths, offsets=offs)
            if st.p12 is None and st.vl is None:
                self._ready[slot].record(self.copy_stream)
        n_real = float(((st.e == 1) & (st.visible != 0)).sum())
        counts = (n_real, float(len(st.lengths)), float(st.visible.sum()))
        return DeviceStep(bags, t, e, vis, st.idx, cid, st.nbytes, counts, slot)

    def __iter__(self) -> Iterator[DeviceStep]:
        it = iter(self.steps)
        pending: List[DeviceStep] = []
        k = 0
        for _ in range(self.depth - 1):
            st = next(it, None)
            if st is None:
                break
            pending.append(self._issue(k % self.depth, st))
            k += 1
        prev: Optional[DeviceStep] = None
        while pending:
            cur = pending.pop(0)
            torch.cuda.current_stream().wait_event(self._ready[cur._slot])
            for ten in (cur.t, cur.e, cur.visible, cur.bags.offsets, cur.cluster_id):
                if ten is not None:
                    ten.record_stream(torch.cuda.current_stream())   # allocated on the copy stream, consumed here
            if prev is not None:
                self._free[prev._slot].record(torch.cuda.current_stream())
            st = next(it, None)
            if st is not None:
                pending.append(self._issue(k % self.depth, st))
                k += 1
            yield cur
            prev = cur
        if prev is not None:
            self._free[prev._slot].record(torch.cuda.current_stream())


def synthetic_steps(n_steps: int, bags_per_step: int, rows_per_bag, C: int = 1024, seed: int = 42, pin: bool = True,
                    event_rate: float = 0.347, labeled_ratio: float = 1.0, distinct: Optional[int] = None,
                    dtype: torch.dtype = torch.float32) -> List[PinnedStep]:
    """Synthetic NLST-shaped steps: randn features (model_stats.py:93), t~U(0,1), e~Bernoulli(0.347).
    rows_per_bag: int or callable(rng) -> int (multiple of 16).  `distinct` < n_steps re-uses buffers cyclically so a long
    run does not need n_steps GiB of host memory."""
    distinct = n_steps if distinct is None else min(distinct, n_steps)
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    made: List[PinnedStep] = []
    for _ in range(distinct):
        lens = [int(rows_per_bag(rng)) if callable(rows_per_bag) else int(rows_per_bag) for _ in range(bags_per_step)]
        bags = [torch.randn(n, C, generator=g) for n in lens]
        labels = [(float(rng.uniform(0.02, 0.98)), float(rng.uniform() < event_rate)) for _ in lens]
        vis = [bool(rng.uniform() < labeled_ratio) for _ in lens]
        made.append(pack_step(bags, labels, visible=vis, pin=pin, dtype=dtype))
    return [made[i % distinct] for i in range(n_steps)]


def loguniform_rows(lo: int = 1024, hi: int = 100000):
    """cfg3 bag sizes: log-uniform in [lo, hi], rounded down to a multiple of 16."""
    def draw(rng):
        n = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        return max(16, n // 16 * 16)
    return draw
