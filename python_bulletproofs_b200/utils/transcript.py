"""Fiat-Shamir transcript: '&'-separated base64 points and decimal scalars
(reference: src/utils/transcript.py:6-33)."""
import base64

from .utils import mod_hash, point_to_b64


class Transcript:
    def __init__(self, seed=b""):
        self.digest = base64.b64encode(seed) + b"&"

    def add_point(self, g):
        self.digest += point_to_b64(g) + b"&"

    def add_list_points(self, gs):
        for g in gs:
            self.add_point(g)

    def add_number(self, x):
        self.digest += str(x).encode() + b"&"

    def get_modp(self, p):
        return mod_hash(self.digest, p)
