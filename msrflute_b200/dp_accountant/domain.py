"""Uniform grid on which privacy-loss distributions are discretised (ref. ``prv_accountant/domain.py``).
``create_aligned`` snaps the grid so that 0 is a grid point and the size is even — required for the FFT
self-composition to stay on the same grid."""
import numpy as np


class Domain:
    def __init__(self, t_min: float, t_max: float, size: int, shifts: float = 0.0):
        if size % 2 != 0:
            raise ValueError("Must have an even size")
        self._t_min, self._t_max, self._size = float(t_min), float(t_max), int(size)
        self._dt = (self._t_max - self._t_min) / (self._size - 1)
        self._shifts = float(shifts)

    @classmethod
    def create_aligned(cls, t_min: float, t_max: float, dt: float) -> "Domain":
        t_min = np.floor(t_min / dt) * dt
        t_max = np.ceil(t_max / dt) * dt
        size = int(np.round((t_max - t_min) / dt)) + 1
        if size % 2 == 1:
            size += 1
            t_max += dt
        d = cls(t_min, t_max, size)
        if np.abs(d.dt() - dt) / dt >= 1e-8:
            raise RuntimeError("grid alignment failed")
        return d

    def shifts(self): return self._shifts
    def size(self): return self._size
    def t_min(self): return self._t_min
    def t_max(self): return self._t_max
    def dt(self): return self._dt
    def t(self, i): return self._t_min + i * self._dt
    def ts(self): return np.linspace(self._t_min, self._t_max, self._size, endpoint=True, dtype=np.longdouble)

    def shift_right(self, dt: float) -> "Domain":
        return Domain(self._t_min + dt, self._t_max + dt, self._size, self._shifts + dt)

    def shift_left(self, dt: float) -> "Domain":
        return self.shift_right(-dt)

    def __eq__(self, o):
        return isinstance(o, Domain) and self._size == o._size and abs(self._t_min - o._t_min) < 1e-12 * max(1, abs(self._t_min)) \
            and abs(self._t_max - o._t_max) < 1e-12 * max(1, abs(self._t_max))

    def __repr__(self):
        return "Domain(t_min={}, t_max={}, size={}, dt={})".format(self._t_min, self._t_max, self._size, self._dt)
