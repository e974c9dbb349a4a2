"""python_bulletproofs_b200 -- B200 (sm_100a) drop-in for the hot path of wborgeaud/python-bulletproofs.

Same Python surface as the reference's `src` package (pippenger / innerproduct / rangeproofs /
utils); the multi-scalar multiplications, generator/scalar folding rounds, scalar-multiplication
batches and verifier equations run in hand-written CUDA behind the C ABI of libbpgpu.so
(include/bp_gpu.h).  No PyTorch, no Triton, no CPU fallback.
"""
from .curve import secp256k1, Curve
from .point import Point

__all__ = ["secp256k1", "Curve", "Point"]
