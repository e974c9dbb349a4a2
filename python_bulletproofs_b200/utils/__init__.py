from .utils import ModP, mod_hash, point_to_bytes, point_to_b64, b64_to_point, bytes_to_point, inner_product, egcd
from .commitments import commitment, vector_commitment
from .transcript import Transcript
from .elliptic_curve_hash import elliptic_hash

__all__ = ["ModP", "mod_hash", "point_to_bytes", "point_to_b64", "b64_to_point", "bytes_to_point", "inner_product",
           "egcd", "commitment", "vector_commitment", "Transcript", "elliptic_hash"]
