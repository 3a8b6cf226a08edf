"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/liboracle.so (dkmc_oracle.c).

The CPU restatement of the DeviceKMC field-and-rate hot path; see dkmc_oracle.h for the
reference citations.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Nothing under devicekmc_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dkmc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_site_dist.restype = C.c_double
        _lib.orc_site_dist.argtypes = [C.c_double] * 6 + [C.c_void_p, C.c_int]
        _lib.orc_rng_uniform.restype = C.c_double
        _lib.orc_select_event.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def site_dist(p1, p2, lattice, pbc):
    lat = _f64(lattice)
    return lib().orc_site_dist(*[float(v) for v in p1], *[float(v) for v in p2], _p(lat), int(pbc))


def neighbor_list(x, y, z, lattice, pbc, nn_dist, method=1):
    """returns (neigh_idx[N, nn], nn)"""
    x, y, z, lat = _f64(x), _f64(y), _f64(z), _f64(lattice)
    N = len(x)
    deg = np.zeros(N, dtype=np.int32)
    nn = lib().orc_neighbor_degrees(N, _p(x), _p(y), _p(z), _p(lat), int(pbc), C.c_double(nn_dist), _p(deg), int(method))
    nb = np.empty((N, nn), dtype=np.int32)
    lib().orc_neighbor_fill(N, _p(x), _p(y), _p(z), _p(lat), int(pbc), C.c_double(nn_dist), nn, _p(nb), int(method))
    return nb, nn


def site_layers(x, start_x, end_x):
    x, s, e = _f64(x), _f64(start_x), _f64(end_x)
    out = np.zeros(len(x), dtype=np.int32)
    lib().orc_site_layers(len(x), _p(x), len(s), _p(s), _p(e), _p(out))
    return out


def update_charge(neigh_idx, element, metals, charge):
    nb, el, me = _i32(neigh_idx), _i32(element), _i32(metals)
    q = _i32(charge).copy()
    N, nn = nb.shape
    lib().orc_update_charge(N, nn, _p(nb), _p(el), _p(me), len(me), _p(q))
    return q


def csr_structure(neigh_idx, NL, NR):
    nb = _i32(neigh_idx)
    N, nn = nb.shape
    m = N - NL - NR
    rp = np.zeros(m + 1, np.int32); lrp = np.zeros(m + 1, np.int32); rrp = np.zeros(m + 1, np.int32)
    lib().orc_csr_row_ptr(N, nn, _p(nb), NL, NR, _p(rp), _p(lrp), _p(rrp))
    ci = np.zeros(max(int(rp[m]), 1), np.int32); lci = np.zeros(max(int(lrp[m]), 1), np.int32); rci = np.zeros(max(int(rrp[m]), 1), np.int32)
    lib().orc_csr_fill(N, nn, _p(nb), NL, NR, _p(rp), _p(ci), _p(lrp), _p(lci), _p(rrp), _p(rci))
    return dict(row_ptr=rp, col=ci[: rp[m]], left_row_ptr=lrp, left_col=lci[: lrp[m]],
                right_row_ptr=rrp, right_col=rci[: rrp[m]])


def assemble_K(neigh_idx, NL, NR, element, charge, metals, high_G, low_G, Vd, row_ptr, col):
    nb, el, q, me = _i32(neigh_idx), _i32(element), _i32(charge), _i32(metals)
    rp, ci = _i32(row_ptr), _i32(col)
    N, nn = nb.shape
    m = N - NL - NR
    val = np.zeros(len(ci), np.float64); rhs = np.zeros(m, np.float64)
    lib().orc_assemble_K(N, nn, _p(nb), NL, NR, _p(el), _p(q), _p(me), len(me), C.c_double(high_G),
                         C.c_double(low_G), C.c_double(Vd), _p(rp), _p(ci), _p(val), _p(rhs))
    return val, rhs


def solve(row_ptr, col, val, rhs, x0=None, tol=1e-13, max_iter=20000, refine=3):
    rp, ci, v, b = _i32(row_ptr), _i32(col), _f64(val), _f64(rhs)
    m = len(b)
    x = np.zeros(m) if x0 is None else _f64(x0).copy()
    info = np.zeros(3)
    lib().orc_solve(m, _p(rp), _p(ci), _p(v), _p(b), _p(x), C.c_double(tol), int(max_iter), int(refine), _p(info))
    return x, info


def background_potential(neigh_idx, NL, NR, element, charge, metals, high_G, low_G, Vd, phi0=None,
                         tol=1e-13, max_iter=20000, refine=3):
    nb, el, q, me = _i32(neigh_idx), _i32(element), _i32(charge), _i32(metals)
    N, nn = nb.shape
    phi = np.zeros(N) if phi0 is None else _f64(phi0).copy()
    info = np.zeros(3)
    lib().orc_background_potential(N, nn, _p(nb), NL, NR, _p(el), _p(q), _p(me), len(me), C.c_double(high_G),
                                   C.c_double(low_G), C.c_double(Vd), _p(phi), C.c_double(tol), int(max_iter),
                                   int(refine), _p(info))
    return phi, info


def laplace_cb_edge(neigh_idx, NL, NR, element, metals, high_G, low_G, Vd, q=1.60217663e-19, phi0=None,
                    tol=1e-13, max_iter=20000, refine=3):
    """Device::setLaplacePotential (potential_solver.cpp:4-139), sparse restatement"""
    nb, el, me = _i32(neigh_idx), _i32(element), _i32(metals)
    N, nn = nb.shape
    phi = np.zeros(N) if phi0 is None else _f64(phi0).copy()
    info = np.zeros(3)
    lib().orc_laplace_cb_edge(N, nn, _p(nb), NL, NR, _p(el), _p(me), len(me), C.c_double(high_G), C.c_double(low_G),
                              C.c_double(Vd), C.c_double(q), _p(phi), C.c_double(tol), int(max_iter), int(refine),
                              _p(info))
    return phi, info


def poisson_gridless(x, y, z, lattice, pbc, charge, sigma, k, rows=None):
    x, y, z, lat, q = _f64(x), _f64(y), _f64(z), _f64(lattice), _i32(charge)
    N = len(x)
    r0, r1 = (0, N) if rows is None else rows
    out = np.zeros(r1 - r0)
    lib().orc_poisson_gridless_rows(N, _p(x), _p(y), _p(z), _p(lat), int(pbc), _p(q), C.c_double(sigma),
                                    C.c_double(k), int(r0), int(r1), _p(out))
    return out


def rate_table(neigh_idx, site_layer, lattice, pbc, T_bg, freq, sigma, k, x, y, z, pb, pc, element, charge, layers):
    """layers: array [n_layers, 4] = E_gen, E_rec, E_Vdiff, E_Odiff"""
    nb, sl, lat = _i32(neigh_idx), _i32(site_layer), _f64(lattice)
    x, y, z, pb, pc = _f64(x), _f64(y), _f64(z), _f64(pb), _f64(pc)
    el, q = _i32(element), _i32(charge)
    L = _f64(layers)
    Eg, Er, Ev, Eo = [np.ascontiguousarray(L[:, c]) for c in range(4)]
    N, nn = nb.shape
    et = np.zeros(N * nn, np.int32); ep = np.zeros(N * nn, np.float64)
    lib().orc_rate_table(N, nn, _p(nb), _p(sl), _p(lat), int(pbc), C.c_double(T_bg), C.c_double(freq),
                         C.c_double(sigma), C.c_double(k), _p(x), _p(y), _p(z), _p(pb), _p(pc), _p(el), _p(q),
                         _p(Eg), _p(Er), _p(Ev), _p(Eo), _p(et), _p(ep))
    return et, ep


class Rng:
    """std::mt19937 + libstdc++ uniform_real_distribution<double>(0,1) (random_num.h:4-23)."""

    class _S(C.Structure):
        _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]

    def __init__(self, seed: int):
        self.s = Rng._S()
        lib().orc_rng_seed(C.byref(self.s), C.c_uint32(seed))

    def uniform(self) -> float:
        return lib().orc_rng_uniform(C.byref(self.s))

    def uniforms(self, n: int) -> np.ndarray:
        return np.array([self.uniform() for _ in range(n)], dtype=np.float64)


def kmc_events(neigh_idx, event_type, event_prob, element, charge, freq, rng: Rng, max_events=65536):
    """runs the residence-time loop; returns (event_time, events[n,4]=(idx,i,j,type), element, charge)"""
    nb = _i32(neigh_idx)
    N, nn = nb.shape
    et, ep = _i32(event_type).copy(), _f64(event_prob).copy()
    el, q = _i32(element).copy(), _i32(charge).copy()
    t = C.c_double(0.0)
    ev = np.zeros(4 * max_events, np.int32)
    ne = lib().orc_kmc_events(N, nn, _p(nb), _p(et), _p(ep), _p(el), _p(q), C.c_double(freq), C.byref(rng.s),
                              C.byref(t), _p(ev), int(max_events))
    return t.value, ev[: 4 * min(ne, max_events)].reshape(-1, 4).copy(), el, q


def set_event_limit(n: int):
    """bench.py only: stop kmc_events after n executed events (0 = no limit)"""
    lib().orc_set_event_limit(int(n))


def select_event(event_prob, u):
    ep = _f64(event_prob)
    ps = C.c_double(0.0)
    idx = lib().orc_select_event(C.c_long(len(ep)), _p(ep), C.c_double(u), C.byref(ps))
    return int(idx), ps.value
