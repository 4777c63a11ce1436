"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU restatement of iw3's stateful depth
normalisers - MinMaxBuffer / EMAMinMaxScaler (iw3/depth_scaler.py:33-142), the `--ema-normalize` look-ahead mode
that SURVEY.md 8f ranks as a "next" row.  The engine implements the stateless default (decay=0, buffer_size=1:
nb200_minmax_map); this file pins the stateful algorithm for the round that ports it.

Pinned against the real reference: tests/golden/depth_scaler.npz (oracle/gen_golden.py depth_scaler).
"""
import numpy as np


def _normalize(frame, mn, mx, mode):
    f = np.float32
    if mode == "minmax":
        scale = f(mx) - f(mn)
        if scale > 0:
            return np.clip((frame - f(mn)) / scale, 0, 1).astype(np.float32)
        return np.clip(frame, 0, 1).astype(np.float32)
    if mx > 0:
        return np.clip(frame / f(mx), 0, 1).astype(np.float32)
    return np.clip(frame, 0, 1).astype(np.float32)


class EMAMinMaxScaler:
    """depth_scaler.py:64-142.  update(frame) -> normalised frame or None while the look-ahead buffer fills."""

    def __init__(self, decay=0.0, buffer_size=1, mode="minmax"):
        assert mode in {"minmax", "max"} and buffer_size > 0
        self.decay, self.buffer_size, self.mode = np.float32(decay), int(buffer_size), mode
        self.reset()

    def reset(self):
        self.min_value = self.max_value = None
        self.queue = []
        self.ring = None       # MinMaxBuffer: 2*buffer_size slots, min/max interleaved, first add fills every slot
        self.count = 0

    def _add(self, mn, mx):
        size = 2 * self.buffer_size
        if self.ring is None:
            self.ring = np.zeros(size, dtype=np.float32)
        if self.count == 0:
            self.ring[0::2] = mn
            self.ring[1::2] = mx
            self.count = 2
        else:
            for v in (mn, mx):
                self.ring[self.count % size] = v
                self.count += 1

    def update(self, frame):
        frame = np.asarray(frame, dtype=np.float32)
        self.queue.append(frame)
        self._add(frame.min(), frame.max())
        if self.count < 2 * self.buffer_size:
            return None
        mn, mx = self.ring.min(), self.ring.max()
        if self.min_value is None:
            self.min_value, self.max_value = mn, mx
        else:
            one = np.float32(1.0)
            self.min_value = self.decay * self.min_value + (one - self.decay) * mn
            self.max_value = self.decay * self.max_value + (one - self.decay) * mx
        return _normalize(self.queue.pop(0), self.min_value, self.max_value, self.mode)

    def flush(self):
        if not self.queue:
            self.reset()
            return []
        if self.min_value is None:
            mn, mx = self.ring.min(), self.ring.max()
        else:
            mn, mx = self.min_value, self.max_value
        out = [_normalize(f, mn, mx, self.mode) for f in self.queue]
        self.reset()
        return out
