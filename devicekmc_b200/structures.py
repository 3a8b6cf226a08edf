"""Synthetic devices for the benchmark configurations.

The base cell is the reference's 2.5 nm TiN/HfO2/Ti/TiN test device (input DATA shipped as
structures/single_devices/test_2.5nm/reordered_device_2.5.xyz: 9 399 sites, lattice
108.97557 x 25.575 x 25.575 A, first/last 144 sites = contact layers), stored as
data/cell_2.5nm.npz.  Larger devices tile it periodically in y and z, which creates no close
pairs (SURVEY.md notes); contacts stay the first/last x-layer and are ordered first/last as the
reference requires (reorder_boundary.py).  Interior sites are ordered by the cell of a uniform
grid (x-major), so that neighbours are close in memory and an x-slab is a contiguous index range.
"""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cell_2.5nm.npz")


def load_base_cell():
    d = np.load(_DATA)
    xyz = d["xyz"]
    return (d["element"].astype(np.int32), xyz[:, 0].copy(), xyz[:, 1].copy(), xyz[:, 2].copy(),
            d["lattice"].copy(), int(d["n_contact"]))


def tile_device(ny: int, nz: int, order: str = "cell", cell_edge: float = 3.6):
    """ny x nz periodic copies of the base cell.  Returns (element, x, y, z, lattice, n_contact).
    order: "cell" (x-major grid-cell order, default), "x" (sorted by x), "tile" (copy after copy,
    each in the reference's own order: lattice atoms first, interstitials last)."""
    el, x, y, z, lat, nc = load_base_cell()
    n = len(x)
    left = np.arange(nc)
    right = np.arange(n - nc, n)
    mid = np.arange(nc, n - nc)
    E, X, Y, Z = [], [], [], []

    def block(idx):
        for iy in range(ny):
            for iz in range(nz):
                E.append(el[idx]); X.append(x[idx]); Y.append(y[idx] + iy * lat[1]); Z.append(z[idx] + iz * lat[2])

    block(left)
    n_left = sum(len(a) for a in E)
    block(mid)
    n_mid_end = sum(len(a) for a in E)
    block(right)
    E, X, Y, Z = (np.concatenate(a) for a in (E, X, Y, Z))
    if order != "tile":
        sl = slice(n_left, n_mid_end)
        if order == "x":
            key = np.argsort(X[sl], kind="stable")
        elif order == "cell":
            cx = np.floor((X[sl] - X.min()) / cell_edge).astype(np.int64)
            cy = np.floor(Y[sl] / cell_edge).astype(np.int64)
            cz = np.floor(Z[sl] / cell_edge).astype(np.int64)
            key = np.lexsort((cz, cy, cx))
        else:
            raise ValueError(order)
        for a in (E, X, Y, Z):
            a[sl] = a[sl][key]
    lattice = np.array([lat[0], lat[1] * ny, lat[2] * nz])
    return E.astype(np.int32), X, Y, Z, lattice, nc * ny * nz


def cell_order(x, y, z, cell_edge: float = 3.6, x0=None):
    """x-major order of sites by the cell of a uniform grid (ties keep the input order): what tile_device's
    "cell" order is, and what Device registers as the solver's internal row order for inputs in the reference's
    order (dkmc_solver_set_order).  Returns the permutation (indices into x)."""
    x0 = float(np.min(x)) if x0 is None else x0
    cx = np.floor((x - x0) / cell_edge).astype(np.int64)
    cy = np.floor(y / cell_edge).astype(np.int64)
    cz = np.floor(z / cell_edge).astype(np.int64)
    return np.lexsort((cz, cy, cx))


def load_structure(name: str):
    """A shipped input structure of the reference in ITS site order (data/<name>.npz, written by
    tools/make_structure_data.py).  Returns (element, x, y, z, lattice, n_contact, V_switch)."""
    d = np.load(os.path.join(os.path.dirname(_DATA), name + ".npz"))
    xyz = d["xyz"]
    return (d["element"].astype(np.int32), xyz[:, 0].copy(), xyz[:, 1].copy(), xyz[:, 2].copy(), d["lattice"].copy(),
            int(d["n_contact"]), float(d["V_switch"]))


def tile_structure(name: str, ny: int, nz: int):
    """ny x nz periodic copies in y, z of a shipped structure (the 20/40 nm crossbars are not in
    the tree, the 10 nm one tiled 2 x 2 stands in).  Every copy keeps the reference's site order; the copies'
    contact sites are ordered first / last as the reference requires."""
    el, x, y, z, lat, nc, vd = load_structure(name)
    n = len(x)
    parts = (np.arange(nc), np.arange(nc, n - nc), np.arange(n - nc, n))
    E, X, Y, Z = [], [], [], []
    for idx in parts:
        for iy in range(ny):
            for iz in range(nz):
                E.append(el[idx]); X.append(x[idx]); Y.append(y[idx] + iy * lat[1]); Z.append(z[idx] + iz * lat[2])
    E, X, Y, Z = (np.concatenate(a) for a in (E, X, Y, Z))
    return E.astype(np.int32), X, Y, Z, np.array([lat[0], lat[1] * ny, lat[2] * nz]), nc * ny * nz, vd


# shipped structures in the reference's site order: name -> (data file, ny, nz)
REAL_WORKLOADS = {
    "device_7.5nm": ("device_7.5nm", 1, 1),                 # BASELINE config 2: 85 071 sites, 10 V constant
    "crossbar_10nm": ("crossbar_10nm_5pitch", 1, 1),        # 110 813 sites, 15 V
    "crossbar_2x2": ("crossbar_10nm_5pitch", 2, 2),         # BASELINE config 3: 443 252 sites, 15 V
}

# named benchmark workloads: (ny, nz)
WORKLOADS = {
    "2.5nm_9k": (1, 1),        # config 0: the shipped test device
    "tiled_100k": (3, 4),      # config 1: ~113 k sites
    "tiled_500k": (7, 8),      # config 2: ~526 k sites
    "tiled_1M": (10, 11),      # config 3: 1 033 890 sites
    "tiled_4M": (21, 20),      # config 4: 3 947 580 sites
}
