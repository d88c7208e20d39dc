"""Import recipe for the live reference (TEST INFRASTRUCTURE, build container only).

/root/reference is read-only and exists only in the build container; on the GPU box the copy staged by
oracle/build_ref.py under oracle/_ref (git-ignored) is used instead.  Its modules import a few
packages that are absent here (matplotlib, h5py, torch_geometric, torch_sparse) at import time only;
empty module stubs are registered so `model.GANSurv`, `model.backbone`, `loss.utils`, `optim`,
`eval.cindex` and `utils.func` import and run on CPU.  Nothing here is used by the
product path or by anything that runs on the GPU box.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (git-ignored, ships via gpurun)


def _resolve_root() -> str:
    env = os.environ.get("ADVMIL_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/model"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns a namespace with the reference modules needed by make_golden.py."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    class _B:  # placeholder base classes
        pass

    if "torch_geometric" not in sys.modules:
        tg = _stub("torch_geometric")
        tg.nn = _stub("torch_geometric.nn", GENConv=object, DeepGCNLayer=object)
        tg.data = _stub("torch_geometric.data", Data=_B, Batch=_B)
    if "torch_sparse" not in sys.modules:
        _stub("torch_sparse", SparseTensor=_B, cat=lambda *a, **k: None)

    import importlib
    ns = types.SimpleNamespace()
    ns.GANSurv = importlib.import_module("model.GANSurv")
    ns.backbone = importlib.import_module("model.backbone")
    ns.backbone_utils = importlib.import_module("model.backbone_utils")
    ns.model_utils = importlib.import_module("model.model_utils")
    ns.loss = importlib.import_module("loss.utils")
    ns.optim = importlib.import_module("optim")
    ns.cindex = importlib.import_module("eval.cindex")
    ns.func = importlib.import_module("utils.func")
    return ns
