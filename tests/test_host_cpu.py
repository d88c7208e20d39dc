"""CPU tests: the C-ABI library loads and exports what include/advmil_b200.h declares (no compute calls), the module
surface matches the reference's state_dict contract, host-side packing/sharding logic, and the data-parallel gradient
maths over a world_size-2 gloo group."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import advmil_oracle as O
from tests.util import build_D, build_G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from advmil_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "advmil_b200.h")).read()
    declared = set(re.findall(r"ADVMIL_API\s+[\w\s\*]+?\b(advmil_\w+)\s*\(", header))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"
    assert lib.advmil_abi_version() == 3
    for i, st in enumerate(_lib.ABI_STRUCTS):
        assert lib.advmil_abi_sizeof(i) == ctypes.sizeof(st)
    assert lib.advmil_gate_packed_width(384) == 768 and lib.advmil_gate_packed_width(128) == 256
    assert lib.advmil_gate_packed_width(32) == 128


def test_product_path_does_not_import_oracle_or_reference():
    for root, _, files in os.walk(os.path.join(ROOT, "advmil_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"
                assert "/root/reference" not in src, f"{f} references the reference tree"


def test_state_dict_contract_matches_reference_names_and_shapes():
    G, D = build_G(device="cpu"), build_D(device="cpu")
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == O.G_SHAPES()
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == O.D_SHAPES()
    assert sum(p.numel() for p in G.parameters()) == 911810
    assert sum(p.numel() for p in D.parameters()) == 206338
    Gc = build_G((1024, 384, 384), mode="cluster", device="cpu")
    assert {k: tuple(v.shape) for k, v in Gc.state_dict().items()} == O.G_CLUSTER_SHAPES()
    from advmil_b200.model.model_utils import init_weights
    G.apply(init_weights)
    assert float(G.backbone.rho[0].bias.abs().max()) == 0.0
    # parameter order handed to the C ABI
    from advmil_b200 import _lib
    assert len(G.gen_params()) == len(_lib.GEN_TENSORS) and len(D.disc_params()) == len(_lib.DISC_TENSORS)


def test_unsupported_configurations_raise_like_the_reference():
    from advmil_b200.model.backbone import load_backbone
    from advmil_b200.model.backbone_utils import make_embedding_layer
    from types import SimpleNamespace
    with pytest.raises(NotImplementedError):
        load_backbone("graph", [1024, 384, 384])
    with pytest.raises(NotImplementedError):
        make_embedding_layer("nope", SimpleNamespace())
    with pytest.raises(AssertionError):
        load_backbone("abmil", [1024, 384])


def test_no_cpu_fallback():
    from advmil_b200 import _lib, ops
    with pytest.raises(_lib.AdvmilError):
        ops.PackedBags(torch.zeros(16, 8), [16])


def test_pack_and_shard_logic():
    from advmil_b200.dataset.packed import group_steps, pack_step, shard_bags_balanced
    bags = [torch.full((n, 8), float(i)) for i, n in enumerate([32, 16, 48])]
    st = pack_step(bags, [(0.1, 1), (0.2, 0), (0.3, 1)], visible=[True, False, True], pin=False)
    assert st.lengths == [32, 16, 48] and st.x.shape == (96, 8)
    assert float(st.x[31, 0]) == 0.0 and float(st.x[32, 0]) == 1.0 and float(st.x[48, 0]) == 2.0
    assert st.visible.tolist() == [1, 0, 1] and st.e.tolist() == [1.0, 0.0, 1.0]
    with pytest.raises(AssertionError):
        pack_step([torch.zeros(40, 8)], [(0.5, 1)], pin=False)          # not a multiple of 16
    assert group_steps(35, 16) == [list(range(0, 16)), list(range(16, 32))]   # trailing partial group dropped
    lens = [100000, 1024, 2048, 50000, 30000, 16, 4096, 70000]
    sh = shard_bags_balanced(lens, 4)
    assert sorted(i for s in sh for i in s) == list(range(8)) and all(len(s) > 0 for s in sh)
    loads = [sum(lens[i] for i in s) for s in sh]
    assert max(loads) == 100000       # the largest bag alone bounds the step; everything else is spread below it
    assert shard_bags_balanced([16, 16], 2) == [[0], [1]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    """Each rank: D-step and G-step gradients of ITS bags with GLOBAL-count loss normalisation, then all-reduce(sum)."""
    import torch.distributed as dist
    from advmil_b200.dataset.packed import shard_bags_balanced
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(1)
    dims, d = (64, 32, 32), 32
    Ns = [48, 160, 96, 32, 208]
    sdG = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(O.G_SHAPES(*dims), 5).items()}
    sdD = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(O.D_SHAPES(64, d, (16, 32)), 6).items()}
    bags = [O.synth_bag(n, 40 + i, 64) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(len(Ns), 3)
    es[1] = 1.0
    vis = [True, True, False, True, True]
    nz = [torch.tensor(np.random.default_rng(50 + i).uniform(size=(1, 16)), dtype=torch.float32) for i in range(len(Ns))]
    mine = shard_bags_balanced(Ns, world)[rank]
    n_real = sum(1 for i in range(len(Ns)) if es[i] == 1 and vis[i])
    n_fake, n_vis = len(Ns), sum(vis)
    # local loss contributions with global denominators (what advmil_disc_loss / advmil_gen_loss compute per rank)
    dl = torch.zeros(())
    gl = torch.zeros(())
    for i in mine:
        x = bags[i]
        with torch.no_grad():
            pred = O.generator_forward(sdG, x, [None, nz[i]], (0, 1))["pred"]
        f_fake = O.prjdisc_forward(sdD, x, pred)["out"].reshape(())
        dl = dl - (1.0 - torch.log(torch.sigmoid(f_fake) + 1e-8)) / n_fake
        if es[i] == 1 and vis[i]:
            f_real = O.prjdisc_forward(sdD, x, ts[i].reshape(1, 1))["out"].reshape(())
            dl = dl - torch.log(torch.sigmoid(f_real) + 1e-8) / n_real
        p2 = O.generator_forward(sdG, x, [None, nz[i]], (0, 1))["pred"]
        ff = O.prjdisc_forward(sdD, x, p2)["out"].reshape(())
        gl = gl + 0.004 * (-ff / n_fake)
        if vis[i]:
            diff = p2.reshape(()) - ts[i]
            gl = gl + (es[i] * diff.abs() + (1 - es[i]) * torch.relu(-diff)) / n_vis
    gD = torch.autograd.grad(dl, list(sdD.values()), allow_unused=True)
    gG = torch.autograd.grad(gl, list(sdG.values()), allow_unused=True)
    flatD = torch.cat([torch.zeros_like(p).reshape(-1) if g_ is None else g_.reshape(-1) for g_, p in zip(gD, sdD.values())])
    flatG = torch.cat([torch.zeros_like(p).reshape(-1) if g_ is None else g_.reshape(-1) for g_, p in zip(gG, sdG.values())])
    dist.all_reduce(flatD)
    dist.all_reduce(flatG)
    if rank == 0:
        q.put((flatD.numpy(), flatG.numpy()))
    dist.destroy_process_group()


def test_data_parallel_gradients_equal_single_process_gloo():
    """N-rank gradients == 1-rank gradients on the concatenated bag list (global-count normalisation)."""
    ctx = mp.get_context("spawn")
    res = {}
    for world in (1, 2):
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res[world] = q.get()
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
    for a, b in zip(res[1], res[2]):
        assert float(np.abs(a - b).max()) <= 1e-6 * max(1.0, float(np.abs(a).max()))
    # and the single-process value is the reference loss' gradient (oracle disc_step_loss with the same inputs)
    dims, d = (64, 32, 32), 32
    Ns = [48, 160, 96, 32, 208]
    sdG = O.synth_state_dict(O.G_SHAPES(*dims), 5)
    sdD = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(O.D_SHAPES(64, d, (16, 32)), 6).items()}
    bags = [O.synth_bag(n, 40 + i, 64) for i, n in enumerate(Ns)]
    ts, es = O.synth_labels(len(Ns), 3)
    es[1] = 1.0
    nz = [[None, torch.tensor(np.random.default_rng(50 + i).uniform(size=(1, 16)), dtype=torch.float32)] for i in range(len(Ns))]
    out = O.disc_step_loss(sdG, sdD, bags, ts, es, [True, True, False, True, True], nz)
    out["loss"].backward()
    ref = torch.cat([torch.zeros_like(p).reshape(-1) if p.grad is None else p.grad.reshape(-1) for p in sdD.values()]).numpy()
    assert float(np.abs(ref - res[2][0]).max()) <= 2e-6 * max(1.0, float(np.abs(ref).max()))


def test_packed_file_round_trip_fp32_and_bf16(tmp_path):
    """On-disk packed format (dataset/packed_file.py): bit-exact round trip in fp32, round-to-nearest-even bf16, ragged bags,
    steps assembled from arbitrary bag indices, and conversion from the reference's per-slide .pt layout."""
    from advmil_b200.dataset.packed import group_steps
    from advmil_b200.dataset.packed_file import PackedFile, pack_reference_layout, write_packed
    g = torch.Generator().manual_seed(5)
    lens = [16, 320, 48, 1600, 16, 96]
    bags = [torch.randn(n, 64, generator=g) for n in lens]
    labels = [(0.1 * i + 0.05, float(i % 2)) for i in range(len(lens))]
    for dt in (torch.float32, torch.bfloat16):
        path = str(tmp_path / f"split_{dt}.advmil")
        info = write_packed(path, iter(bags), labels, dtype=dt, names=[f"pat{i}" for i in range(len(lens))])
        assert info["rows"] == sum(lens) and info["C"] == 64
        pf = PackedFile(path)
        assert len(pf) == len(lens) and pf.lengths == lens and pf.names[3] == "pat3" and pf.dtype == dt
        assert np.array_equal(pf.labels, np.asarray(labels, dtype=np.float32))
        for i, b in enumerate(bags):
            assert torch.equal(pf.bag(i), b.to(dt))
        st = pf.step([3, 0, 5], visible=[True, False, True], pin=False)
        assert st.lengths == [1600, 16, 96] and st.x.dtype == dt
        assert torch.equal(st.x, torch.cat([bags[3], bags[0], bags[5]]).to(dt))
        assert st.offsets.tolist() == [0, 1600, 1616, 1712] and st.visible.tolist() == [1, 0, 1]
        assert torch.allclose(st.t, torch.tensor([labels[3][0], labels[0][0], labels[5][0]]))
        assert st.nbytes == st.x.numel() * st.x.element_size() + 3 * 8 + 3
    assert group_steps(len(lens), 4) == [[0, 1, 2, 3]]      # the trailing partial group is dropped (model_handler.py:321)
    with pytest.raises(AssertionError):
        write_packed(str(tmp_path / "bad.advmil"), iter([torch.randn(17, 8)]), [(0.5, 1.0)])
    with pytest.raises(ValueError):
        open(tmp_path / "junk.advmil", "wb").write(b"x" * 128)
        PackedFile(str(tmp_path / "junk.advmil"))
    # reference layout: one .pt per slide, several slides per patient, concatenated in order; ragged tail trimmed to 16
    slides = {"p0": [], "p1": []}
    for pid, ns in (("p0", [40, 24]), ("p1", [35])):
        for k, n in enumerate(ns):
            f = str(tmp_path / f"{pid}_{k}.pt")
            torch.save(torch.randn(n, 32, generator=g), f)
            slides[pid].append(f)
    info = pack_reference_layout(slides, {"p0": (0.3, 1.0), "p1": (0.8, 0.0)}, str(tmp_path / "ref.advmil"))
    pf = PackedFile(str(tmp_path / "ref.advmil"))
    assert pf.lengths == [64, 32] and pf.names == ["p0", "p1"]
    assert torch.equal(pf.bag(0), torch.cat([torch.load(f) for f in slides["p0"]]))
    assert torch.equal(pf.bag(1), torch.load(slides["p1"][0])[:32])


def test_p12_transport_format_round_trip_is_bit_exact():
    """dataset/codec.py: the 12-bit transport form of bf16 features decodes to the same 16-bit words for every kind of
    value (zeros of both signs, denormals, infinities, NaN) and for matrices with more than 15 distinct exponents
    (escapes); Gaussian features cost 12.0x bits per element."""
    from advmil_b200.dataset.codec import decode_p12_host, encode_bf16_p12
    g = torch.Generator().manual_seed(5)
    x = torch.randn(512, 1024, generator=g).to(torch.bfloat16)
    x[0, :8] = torch.tensor([0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1e-40, -3e38, 1.0]).to(torch.bfloat16)
    x[5] = (torch.randn(1024, generator=g) * torch.logspace(-30, 30, 1024)).to(torch.bfloat16)
    p = encode_bf16_p12(x)
    assert p.esc_idx.numel() > 0 and p.lo.numel() == x.numel() and p.hi.numel() == x.numel() // 2
    assert torch.equal(decode_p12_host(p).view(torch.int16), x.view(torch.int16))
    y = torch.randn(256, 1024, generator=g).to(torch.bfloat16)
    q = encode_bf16_p12(y)
    assert q.nbytes / (y.numel() * 2) < 0.7505 and torch.equal(decode_p12_host(q).view(torch.int16), y.view(torch.int16))
    z = torch.relu(torch.randn(64, 1024, generator=g)).to(torch.bfloat16)          # post-ReLU features: zeros take a table entry
    assert torch.equal(decode_p12_host(encode_bf16_p12(z)).view(torch.int16), z.view(torch.int16))
    # degenerate inputs: the smallest block, a constant matrix (one table entry, no escapes), every exponent present
    for w in (torch.randn(8, generator=g).to(torch.bfloat16), torch.full((16, 16), 3.0, dtype=torch.bfloat16),
              torch.arange(0, 65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).reshape(64, 1024)):
        pw = encode_bf16_p12(w)
        assert torch.equal(decode_p12_host(pw).view(torch.int16), w.view(torch.int16))
    assert encode_bf16_p12(torch.full((16, 16), 3.0, dtype=torch.bfloat16)).esc_idx.numel() == 0


def test_packed_file_p12_storage_round_trip(tmp_path):
    """write_packed(transport="p12"): one code table for the file, planes of any bag selection concatenate into a step;
    bag() and step() reproduce the bf16 words of the raw bf16 file exactly, at 3/4 of its feature bytes."""
    from advmil_b200.dataset.codec import decode_p12_host
    from advmil_b200.dataset.packed_file import PackedFile, write_packed
    g = torch.Generator().manual_seed(11)
    lens = [64, 320, 16, 160, 96]
    bags = [torch.randn(n, 1024, generator=g) for n in lens]
    bags[2][3] = torch.randn(1024, generator=g) * torch.logspace(-30, 30, 1024)        # escapes inside the smallest bag
    labels = [(0.1 * i, float(i % 2)) for i in range(len(lens))]
    raw, p12 = str(tmp_path / "raw.advmil"), str(tmp_path / "p12.advmil")
    write_packed(raw, iter(bags), labels, dtype=torch.bfloat16)
    info = write_packed(p12, iter(bags), labels, dtype=torch.bfloat16, transport="p12")
    assert info["escapes"] > 0
    fr, fp = PackedFile(raw), PackedFile(p12)
    assert fp.lengths == lens and fp.dtype == torch.bfloat16 and np.array_equal(fp.labels, fr.labels)
    for i in range(len(lens)):
        assert torch.equal(fp.bag(i).view(torch.int16), fr.bag(i).view(torch.int16))
    sel = [3, 2, 0]
    sp, sr = fp.step(sel, pin=False), fr.step(sel, pin=False)
    assert sp.p12 is not None and sp.lengths == sr.lengths and sp.nbytes < 0.78 * sr.nbytes
    assert torch.equal(decode_p12_host(sp.p12).view(torch.int16), sr.x.view(torch.int16))
    feats = lambda path: os.path.getsize(path) - PackedFile(path)._feats_pos                  # noqa: E731
    assert feats(p12) < 0.77 * feats(raw)


def test_vl_transport_format_round_trip_is_bit_exact():
    """The entropy-coded transport form (host encoder + the library's sequential reference decoder, no GPU): exact for
    Gaussian features, for every bf16 bit pattern (escapes), for a constant matrix (one 1-bit code) and a minimal block;
    ~10.9 bits per element on Gaussian features."""
    from advmil_b200.dataset.codec import decode_vl_host, encode_bf16_vl
    torch.manual_seed(3)
    x = torch.randn(64 * 4096).to(torch.bfloat16)
    p = encode_bf16_vl(x)
    assert torch.equal(decode_vl_host(p).view(torch.int16), x.view(torch.int16))
    assert 10.5 < p.bits_per_element < 11.1 and int(p.tab_len.max()) <= 8
    kraft = sum(2.0 ** -int(l) for l in p.tab_len if l)
    assert abs(kraft - 1.0) < 1e-12
    allbits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).repeat(2)
    pa = encode_bf16_vl(allbits)
    assert pa.esc_idx.numel() > 0 and torch.equal(decode_vl_host(pa).view(torch.int16), allbits.view(torch.int16))
    c = torch.full((4096,), -0.375).to(torch.bfloat16)
    pc = encode_bf16_vl(c)
    assert torch.equal(decode_vl_host(pc).view(torch.int16), c.view(torch.int16)) and pc.bits_per_element < 9.5
    nn_ = torch.relu(torch.randn(8 * 4096)).mul(0.5).to(torch.bfloat16)          # non-negative features with exact zeros
    assert torch.equal(decode_vl_host(encode_bf16_vl(nn_)).view(torch.int16), nn_.view(torch.int16))


def test_vl_packed_file_round_trip_and_step_assembly(tmp_path):
    """transport="vl" on disk: one Huffman table per file, every bag encoded on its own; single bags and arbitrary step
    selections decode (host reference decoder) to the bf16 words of the raw file, special values included; the file is
    smaller than the p12 file."""
    import os
    from advmil_b200.dataset.codec import decode_vl_host
    from advmil_b200.dataset.packed_file import PackedFile, write_packed
    g = torch.Generator().manual_seed(17)
    lens = [320, 1600, 48, 16, 2048]
    bags = [torch.randn(n, 1024, generator=g) for n in lens]
    bags[2][0, :7] = 0.0
    bags[3].view(-1)[5] = float("inf")
    bags[3].view(-1)[9] = 1e-40
    labels = [(0.1 * i, float(i % 2)) for i in range(len(lens))]
    vl = write_packed(str(tmp_path / "a.vl"), iter(bags), labels, dtype=torch.bfloat16, transport="vl", names=[f"p{i}" for i in range(5)])
    p12 = write_packed(str(tmp_path / "a.p12"), iter(bags), labels, dtype=torch.bfloat16, transport="p12")
    assert vl["bytes"] < 0.93 * p12["bytes"] and vl["bits_per_element"] < 11.1
    pf = PackedFile(str(tmp_path / "a.vl"))
    assert pf.names == [f"p{i}" for i in range(5)] and pf.lengths == lens
    for i, b in enumerate(bags):
        assert torch.equal(pf.bag(i).view(torch.int16), b.to(torch.bfloat16).view(torch.int16))
    st = pf.step([3, 1, 4], pin=False)
    want = torch.cat([bags[i] for i in (3, 1, 4)]).to(torch.bfloat16)
    assert torch.equal(decode_vl_host(st.vl).view(torch.int16), want.view(torch.int16))
    assert st.lengths == [16, 1600, 2048] and st.nbytes < 0.70 * want.numel() * 2


def test_split_tf32_arithmetic_is_fp32_grade():
    """Numerics of the split-tf32 ("3xTF32") contractions (csrc/rlip_chain.cu split3 / mma_chunk, csrc/gemm_tc.cu X3), restated
    in numpy: hi = rna_tf32(x) (integer add + mask, as the kernels do it), lo = x - hi exactly, the tensor core sees lo
    truncated to tf32, a product is hi.hi + lo.hi + hi.lo with the small terms summed apart from the main chain.  The
    dropped lo.lo term and the truncation of lo bound the relative error of a product by ~2^-20; a K = 128 dot product
    (the region-level chain of the RLIP head, reference model/model_utils.py:202-210) stays within 1e-6 of sum |a||b|."""
    rng = np.random.default_rng(0)

    def split(x):
        bits = x.view(np.uint32)
        hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
        lo = (x - hi).astype(np.float32)                                  # exact: |lo| <= 2^-11 |x|
        assert np.array_equal(lo.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))
        lo_t = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)   # what the tensor core reads
        return hi, lo_t

    a = (rng.standard_normal((4096, 128)) * np.exp(rng.uniform(-3, 3, (4096, 128)))).astype(np.float32)
    b = (rng.standard_normal((4096, 128)) * np.exp(rng.uniform(-3, 3, (4096, 128)))).astype(np.float32)
    ah, al = split(a)
    bh, bl = split(b)
    assert float(np.abs(a - ah).max() / np.abs(a).max()) < 2.0 ** -10    # plain tf32 would stop here: ~5e-4 per operand
    exact = a.astype(np.float64) * b.astype(np.float64)
    main = ah.astype(np.float64) * bh.astype(np.float64)
    small = al.astype(np.float64) * bh.astype(np.float64) + ah.astype(np.float64) * bl.astype(np.float64)
    rel = np.abs(main + small - exact) / np.abs(exact)
    assert float(rel.max()) < 2.0 ** -19.5, float(rel.max())
    # dot products with fp32 accumulation of the two chains (round-to-nearest here; the tensor core truncates, which is why the
    # kernels keep the small terms out of the main accumulator and add the two once at the end)
    dot = (main.astype(np.float32).sum(1, dtype=np.float32) + small.astype(np.float32).sum(1, dtype=np.float32)).astype(np.float64)
    ref = exact.sum(1)
    bound = (np.abs(a).astype(np.float64) * np.abs(b).astype(np.float64)).sum(1)
    assert float((np.abs(dot - ref) / bound).max()) < 1e-6
    # plain tf32 (hi.hi only) misses the 1e-5 parity bar of the fp32 mode by two orders of magnitude
    assert float((np.abs(main.sum(1) - ref) / bound).max()) > 1e-5
