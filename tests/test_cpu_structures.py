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


def test_internal_row_order_makes_an_index_range_a_slab():
    """the slab plan of the multi-GPU solver on the shipped 7.5 nm device: in the reference's site order a rank's rows
    touch columns all over the matrix (the halo degenerates to an all-gather); in the solver's internal x-major
    cell order (what dkmc_solver_set_order registers) every rank needs one narrow band from each x-neighbour"""
    from oracle import oracle as O
    from devicekmc_b200 import slab
    from devicekmc_b200.host import KMCParameters
    el, x, y, z, lat, nc, vd = S.load_structure("device_7.5nm")
    p = KMCParameters(lattice=tuple(lat), num_atoms_contact=nc, num_atoms_first_layer=nc)
    nb, nn = O.neighbor_list(x, y, z, lat, p.pbc, p.nn_dist, method=1)
    cs = O.csr_structure(nb, nc, nc)
    rp, col = cs["row_ptr"].astype(np.int64), cs["col"].astype(np.int64)
    m = len(rp) - 1
    world, tile = 4, 1984

    def halo_rows(rp_, col_):
        rows = slab.split_rows_by_nnz(rp_, world, tile)
        halos = slab.halo_ranges(rp_, col_, rows)
        recv, send = slab.halo_plan(rows, halos)
        return rows, sum(e - b for r in recv for (_, b, e) in r), max(len(r) for r in recv)
    rows_ref, halo_ref, peers_ref = halo_rows(rp, col)
    # P K P^T: rows in the internal order, columns renumbered and sorted (what order_fill_kernel builds)
    sl = slice(nc, len(x) - nc)
    order = S.cell_order(x[sl], y[sl], z[sl], x0=float(x.min()))
    inv = np.empty(m, np.int64); inv[order] = np.arange(m)
    lens = (rp[1:] - rp[:-1])[order]
    rp2 = np.concatenate([[0], np.cumsum(lens)])
    col2 = np.empty_like(col)
    for pos, r in enumerate(order):
        col2[rp2[pos]:rp2[pos + 1]] = np.sort(inv[col[rp[r]:rp[r + 1]]])
    rows_int, halo_int, peers_int = halo_rows(rp2, col2)
    assert halo_ref > 2.5 * m and peers_ref == world - 1   # reference order: every rank receives most of the vector from everybody
    assert peers_int <= 2                                  # internal order: the two x-neighbours only,
    assert halo_int < 0.4 * m                              # one 3.6 A layer of cells per face and direction (this device is
    assert halo_int < halo_ref / 6                         # only ~17 cells long in x, so a layer is 6 % of the rows)
