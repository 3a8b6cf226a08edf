"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/_ref/libdkmc_ref.so.

libdkmc_ref.so is the UNMODIFIED reference CPU build (manasakani/DeviceKMC) plus our
harness `ref_harness.cpp`; see oracle/Makefile.  It is the live oracle used to pin
`dkmc_oracle.c` and to generate the fixtures under tests/golden/.  Nothing in the
product path (`devicekmc_b200/`) imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libdkmc_ref.so")


def available() -> bool:
    return os.path.exists(REF_LIB)


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_LIB)
        _lib.ref_create.restype = C.c_void_p
        _lib.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        _lib.ref_site_dist.restype = C.c_double
        _lib.ref_site_dist.argtypes = [C.c_double] * 6 + [C.c_void_p, C.c_int]
        _lib.ref_get_layers.restype = C.c_int
        _lib.ref_kmc_step.restype = C.c_int
        _lib.ref_background_potential.argtypes = [C.c_void_p, C.c_double, C.c_int]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def site_dist(p1, p2, lattice, pbc) -> float:
    lib = _load()
    lat = np.ascontiguousarray(lattice, dtype=np.float64)
    return lib.ref_site_dist(*[float(v) for v in p1], *[float(v) for v in p2], _p(lat), int(pbc))


class RefSim:
    """One reference Device + KMCProcess pair (reference seeds: rnd_seed from the
    parameters file for the device, rnd_seed_kmc = 1 for the KMC stream)."""

    def __init__(self, params: str, xyz_a: str, xyz_b: str | None = None, substoich: bool = True):
        lib = _load()
        self._lib = lib
        self._h = C.c_void_p(
            lib.ref_create(params.encode(), xyz_a.encode(), (xyz_b or "").encode(), int(substoich))
        )
        ii = np.zeros(8, dtype=np.int32)
        lib.ref_info_int(self._h, _p(ii))
        (self.N, self.nn, self.N_atom, self.num_atoms_contact, self.num_atoms_first_layer,
         self.pbc, self.num_metals, self.rnd_seed) = [int(v) for v in ii]
        dd = np.zeros(10, dtype=np.float64)
        lib.ref_info_double(self._h, _p(dd))
        self.sigma, self.k, self.T_bg, self.freq, self.nn_dist, self.high_G, self.low_G = [float(v) for v in dd[:7]]
        self.lattice = dd[7:10].copy()
        m = np.zeros(self.num_metals, dtype=np.int32)
        lib.ref_get_metals(self._h, _p(m))
        self.metals = m
        lay = np.zeros(4 * 16, dtype=np.float64)
        nl = lib.ref_get_layers(self._h, _p(lay))
        self.layers = lay[: 4 * nl].reshape(nl, 4).copy()  # E_gen, E_rec, E_Vdiff, E_Odiff

    def close(self):
        if self._h:
            self._lib.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters / setters -------------------------------------------------
    def _get(self, fn, n, dtype):
        out = np.zeros(n, dtype=dtype)
        getattr(self._lib, fn)(self._h, _p(out))
        return out

    def positions(self):
        x = np.zeros(self.N); y = np.zeros(self.N); z = np.zeros(self.N)
        self._lib.ref_get_positions(self._h, _p(x), _p(y), _p(z))
        return x, y, z

    def neigh_idx(self):
        return self._get("ref_get_neigh_idx", self.N * self.nn, np.int32).reshape(self.N, self.nn)

    def site_layer(self):
        return self._get("ref_get_site_layer", self.N, np.int32)

    def element(self):
        return self._get("ref_get_element", self.N, np.int32)

    def charge(self):
        return self._get("ref_get_charge", self.N, np.int32)

    def potential_boundary(self):
        return self._get("ref_get_potential_boundary", self.N, np.float64)

    def potential_charge(self):
        return self._get("ref_get_potential_charge", self.N, np.float64)

    def set_element(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32); self._lib.ref_set_element(self._h, _p(a))

    def set_charge(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32); self._lib.ref_set_charge(self._h, _p(a))

    def set_potential_boundary(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64); self._lib.ref_set_potential_boundary(self._h, _p(a))

    def set_potential_charge(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64); self._lib.ref_set_potential_charge(self._h, _p(a))

    # ---- the path ----------------------------------------------------------
    def update_charge(self):
        self._lib.ref_update_charge(self._h)

    def background_potential(self, Vd: float, n_contact: int = 0):
        self._lib.ref_background_potential(self._h, float(Vd), int(n_contact))

    def laplace_cb_edge(self, Vd: float):
        """Device::setLaplacePotential (CPU branch): returns site_CB_edge"""
        self._lib.ref_laplace_cb_edge.argtypes = [C.c_void_p, C.c_double]
        self._lib.ref_laplace_cb_edge(self._h, float(Vd))
        return self._get("ref_get_cb_edge", self.N, np.float64)

    def write_snapshot(self, filename: str, folder: str):
        """Device::writeSnapshot: writes ./<folder>/<filename> relative to the current directory"""
        self._lib.ref_write_snapshot.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        self._lib.ref_write_snapshot(self._h, filename.encode(), folder.encode())

    def poisson_gridless(self):
        self._lib.ref_poisson_gridless(self._h)

    def rate_table(self):
        n = self.N * self.nn
        et = np.zeros(n, dtype=np.int32); ep = np.zeros(n, dtype=np.float64)
        self._lib.ref_rate_table(self._h, _p(et), _p(ep))
        return et, ep

    def kmc_step(self, max_events: int = 4096):
        t = C.c_double(0.0)
        ev = np.zeros(2 * max_events, dtype=np.int32)
        ne = self._lib.ref_kmc_step(self._h, C.byref(t), _p(ev), max_events)
        return t.value, ev[: 2 * min(ne, max_events)].reshape(-1, 2).copy()

    def peek_kmc_rng(self, n: int):
        out = np.zeros(n, dtype=np.float64)
        self._lib.ref_peek_kmc_rng(self._h, _p(out), n)
        return out
