"""Host-side mirror of `zaru::timer::Timer` (crates/zaru/src/timer.rs:16-97).

The reference measures wall-clock `Instant` deltas around its CPU stages; here the samples are the DEVICE
milliseconds the library reports per call (`zb_detector_timers`, `zb_estimator_timers`).  Averaging (EMA with
alpha 0.3), the sample count and the reset-on-display behaviour are the reference's.
"""
from __future__ import annotations

import numpy as np

EMA_ALPHA = np.float32(0.3)


class Timer:
    def __init__(self, name: str):
        self.name = name
        self._has = False
        self._avg = np.float32(0.0)
        self._count = 0

    def record(self, seconds: float):
        """`Timer::stop`: filter the duration (seconds, f32) through the EMA (filter/ema.rs:29-42)."""
        x = np.float32(seconds)
        if self._has:
            self._avg = EMA_ALPHA * x + (np.float32(1.0) - EMA_ALPHA) * self._avg
        else:
            self._has, self._avg = True, x
        self._count += 1

    def __str__(self):
        """`Display`: prints the average and resets (timer.rs:76-88)."""
        avg_ms, n = float(self._avg) * 1000.0, self._count
        self._has, self._avg, self._count = False, np.float32(0.0), 0
        return f"{self.name}: {n}x{avg_ms:.1f}ms"
