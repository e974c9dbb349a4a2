"""Device-resident operand vectors (hard part 2): marshal once, reuse."""
import ctypes

from . import _native as nat


class _Handle:
    _upload = None

    def __init__(self, raw, n):
        h = ctypes.c_uint64()
        nat.check(getattr(nat.load(), self._upload)(raw, n, ctypes.byref(h)))
        self.handle, self.n = h, n

    def __len__(self):
        return self.n

    def free(self):
        if self.handle is not None:
            nat.load().bp_handle_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:   # interpreter shutdown
            pass


class DevicePoints(_Handle):
    _upload = "bp_points_upload"

    def __init__(self, points=None, raw=None):
        if raw is None:
            raw = nat.pack_points(points)
        super().__init__(raw, len(raw) // 64)


    def precompute(self, window_bits=0):
        """Store the window multiples 2^(c*w) * P_i next to the vector (bp_points_precompute): later multiexps over this handle
        run without doublings.  Worth it for vectors that are used again and again (generators)."""
        nat.check(nat.load().bp_points_precompute(self.handle, window_bits))
        return self


class DeviceScalars(_Handle):
    _upload = "bp_scalars_upload"

    def __init__(self, scalars=None, raw=None):
        if raw is None:
            raw = nat.pack_scalars(scalars)
        super().__init__(raw, len(raw) // 32)
