"""Host-side mirror of ``src/simulation/state.rs`` over the CUDA library: ``State<T>`` whose flat
state vector lives on the device (``fp_state_*``), for the reference's ``Stateful`` types, and the
SPH neighbour pass (``fp_sph_neighbors``).

The reference's ``State`` is a value that lives one simulation step (``State::new(elements)``,
``euler_step`` / ``rk4_step`` return a NEW ``State``, ``get_elements`` consumes it); here the
vector stays resident and the step methods advance it in place -- ``as_vector`` /
``get_elements`` read it back."""
from __future__ import annotations

import ctypes as C
from enum import Enum

import numpy as np

from . import _lib
from ._lib import check, f32c, ptr


class Integration(Enum):
    """``state::Integration`` (state.rs:4-8)"""
    Euler = 0
    Rk4 = 1


class State:
    """``State<T>`` for ``T`` one of ``_lib.STATEFUL_*`` (state.rs:37-113)."""

    def __init__(self, kind: int, elements, device: int = 0):
        self._lib = _lib.load()
        self.kind = kind
        self.k = self._lib.fp_state_num_state_elements(kind)
        if not self.k:
            raise ValueError("unknown Stateful kind")
        st = f32c(elements, (-1, self.k))   # panics in the reference: "State Vector incorrect size!"
        self.n = len(st)
        self._h = C.c_void_p()
        check(self._lib.fp_state_create(C.byref(self._h), device, kind, self.n, ptr(st) if self.n else None))

    new = classmethod(lambda cls, kind, elements, **kw: cls(kind, elements, **kw))

    @classmethod
    def from_state_vector(cls, kind: int, state_vector, **kw) -> "State":
        return cls(kind, state_vector, **kw)

    def num_state_elements(self) -> int:
        return self.k

    def derivative(self) -> np.ndarray:
        out = np.empty(self.n * self.k, np.float32)
        check(self._lib.fp_state_derivative(self._h, ptr(out)))
        return out

    def as_vector(self) -> np.ndarray:
        out = np.empty(self.n * self.k, np.float32)
        check(self._lib.fp_state_read(self._h, ptr(out)))
        return out

    def get_elements(self) -> np.ndarray:
        return self.as_vector().reshape(self.n, self.k)

    def euler_step(self, timestep: float, nsteps: int = 1) -> "State":
        check(self._lib.fp_state_euler_step(self._h, float(np.float32(timestep)), nsteps))
        return self

    def rk4_step(self, timestep: float, nsteps: int = 1) -> "State":
        check(self._lib.fp_state_rk4_step(self._h, float(np.float32(timestep)), nsteps))
        return self

    def step(self, integration: Integration, timestep: float) -> "State":
        """springy::Simulation::step's ``match self.config.integration`` (springy/simulation.rs:33-36)"""
        return self.rk4_step(timestep) if integration == Integration.Rk4 else self.euler_step(timestep)

    def time_steps(self, timestep: float, rk4: bool, launches: int) -> float:
        """-> milliseconds (CUDA events) for ``launches`` single-step passes"""
        ms = C.c_float(0)
        check(self._lib.fp_state_time_steps(self._h, float(np.float32(timestep)), 1 if rk4 else 0, launches,
                                            C.byref(ms)))
        return ms.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fp_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sph_neighbors(positions, k: int = 8, kernal_max_distance: float = 0.1, particle_mass: float = 0.001,
                  device: int = 0):
    """The neighbour pass of ``sph::Simulation::step`` (sph/mod.rs:89-121): for every particle its
    ``k`` nearest (itself included, ascending distance) within ``kernal_max_distance`` and the
    Monaghan density over them.  -> (index [n, k] (0xffffffff past the count), count [n], density [n])"""
    lib = _lib.load()
    p = f32c(positions, (-1, 3))
    n = len(p)
    idx = np.empty((n, k), np.uint32)
    cnt = np.empty(n, np.uint32)
    den = np.empty(n, np.float32)
    ms = C.c_float(0)
    check(lib.fp_sph_neighbors(device, n, ptr(p), k, float(np.float32(kernal_max_distance)),
                               float(np.float32(particle_mass)), ptr(idx), ptr(cnt), ptr(den), C.byref(ms)))
    return idx, cnt, den
