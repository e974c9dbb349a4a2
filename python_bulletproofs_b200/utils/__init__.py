from .utils import (ModP, mod_hash, point_to_bytes, point_to_b64, b64_to_point, bytes_to_point, bytes_to_points,
                    b64_to_points, inner_product, egcd)
from .commitments import commitment, vector_commitment
from .transcript import Transcript
from .elliptic_curve_hash import elliptic_hash, elliptic_hash_batch

__all__ = ["ModP", "mod_hash", "point_to_bytes", "point_to_b64", "b64_to_point", "bytes_to_point", "bytes_to_points", "b64_to_points", "elliptic_hash_batch", "inner_product",
           "egcd", "commitment", "vector_commitment", "Transcript", "elliptic_hash"]
