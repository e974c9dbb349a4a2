"""src/pippenger/__init__.py:1-7 -- the module singleton every caller goes through."""
from ..curve import secp256k1
from .pippenger import Pippenger
from .group import EC, Group

PipSECP256k1 = Pippenger(EC(secp256k1))

__all__ = ["Pippenger", "EC", "Group", "PipSECP256k1"]
