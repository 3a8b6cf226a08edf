"""Host-side structure handling (no GPU): the shipped structures' data files, the tiling of a structure, and
the x-major cell order the solver uses internally for inputs in the reference's order."""
import numpy as np

from devicekmc_b200 import structures as S


def test_shipped_structures_load_in_reference_order():
    el, x, y, z, lat, nc, vd = S.load_structure("device_7.5nm")
    assert len(x) == 85071 and nc == 12960 and vd == 10.0
    # the contact layers (ten slices of Ti / N at either end) first / last
    assert np.all(np.isin(el[:nc], (6, 8))) and np.all(np.isin(el[-nc:], (6, 8)))
    assert x[:nc].max() < x[nc:-nc].mean() < x[-nc:].min()
    el, x, y, z, lat, nc, vd = S.load_structure("crossbar_10nm_5pitch")
    assert len(x) == 110813 and nc == 11520 and vd == 15.0


def test_tile_structure_keeps_contacts_first_and_last():
    el, x, y, z, lat, nc, vd = S.tile_structure("crossbar_10nm_5pitch", 2, 2)
    e1, x1, y1, z1, lat1, nc1, _ = S.load_structure("crossbar_10nm_5pitch")
    assert len(x) == 4 * len(x1) and nc == 4 * nc1
    assert np.allclose(lat, [lat1[0], 2 * lat1[1], 2 * lat1[2]])
    assert x[:nc].max() == x1[:nc1].max() and x[-nc:].min() == x1[-nc1:].min()
    # the four copies of the first contact site
    assert sorted(zip(y[:nc:nc1] - y1[0], z[:nc:nc1] - z1[0])) == [(0.0, 0.0), (0.0, lat1[2]), (lat1[1], 0.0), (lat1[1], lat1[2])]
    assert np.array_equal(np.bincount(el, minlength=10), 4 * np.bincount(e1, minlength=10))


def test_cell_order_is_identity_on_the_synthetic_tiles_and_a_permutation_otherwise():
    el, x, y, z, lat, nc = S.tile_device(2, 2)
    sl = slice(nc, len(x) - nc)
    assert np.array_equal(S.cell_order(x[sl], y[sl], z[sl], x0=float(x.min())), np.arange(len(x) - 2 * nc))
    el, x, y, z, lat, nc, vd = S.load_structure("device_7.5nm")
    sl = slice(nc, len(x) - nc)
    o = S.cell_order(x[sl], y[sl], z[sl], x0=float(x.min()))
    assert np.array_equal(np.sort(o), np.arange(len(o))) and not np.array_equal(o, np.arange(len(o)))
    # x-major: the cell index along x never decreases in the new order
    cx = np.floor((x[sl][o] - x.min()) / 3.6)
    assert np.all(np.diff(cx) >= 0)
