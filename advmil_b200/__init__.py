"""advmil_b200 — B200-native (sm_100a) implementation of the AdvMIL adversarial G+D hot path.

Module surface mirrors the reference (liupei101/AdvMIL): `advmil_b200.model.GANSurv.{Generator, PrjDiscriminator}`,
`advmil_b200.model.backbone.load_backbone`, `advmil_b200.model.model_utils.init_weights`, so that
model/model_handler.py runs unchanged after swapping its four model imports.
All arithmetic runs in libadvmil_b200.so (hand-written CUDA, C ABI in include/advmil_b200.h).
"""
import os

_PRECISION = os.environ.get("ADVMIL_PRECISION", "fp32")


def set_precision(mode: str) -> None:
    """'fp32' (FFMA, exact-fp32 parity), 'tf32' (tcgen05 kind::tf32 on the fp32 data as stored) or 'bf16' (bf16 storage of
    x and the [rows, *] activations, tcgen05 kind::f16, fp32 accumulation / statistics / parameters)."""
    global _PRECISION
    assert mode in ("fp32", "tf32", "tf32x3", "bf16")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION
