"""Multi-rank host logic on the CPU (gloo, world_size 2 and 3): the slab plan of
devicekmc_b200/slab.py — row ranges by SpMV tiles, halo segments, all-gather chunks, rate-sum
owner — drives a numpy emulation of the distributed PCG whose result must equal the
single-process solve.  The arithmetic here is test scaffolding; the planning code is the product's."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _system():
    """K of config 0 at Vd = 2 V through the oracle (test infrastructure)"""
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters
    from oracle import oracle as O
    el, x, y, z, lat, nc = S.load_base_cell()
    p = KMCParameters()
    nb, nn = O.neighbor_list(x, y, z, lat, 0, p.nn_dist)
    el = el.copy(); el[(el == 3) & (np.arange(len(el)) % 17 == 0)] = 2      # some vacancies
    q = O.update_charge(nb, el, p.metals, np.zeros(len(el), np.int32))
    csr = O.csr_structure(nb, nc, nc)
    val, rhs = O.assemble_K(nb, nc, nc, el, q, p.metals, p.high_G, p.low_G, 2.0, csr["row_ptr"], csr["col"])
    return csr["row_ptr"], csr["col"], val, rhs


def _worker(rank, world, port, tile, out_dir):
    import torch
    import torch.distributed as dist
    from devicekmc_b200 import slab
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    row_ptr, col, val, rhs = _system()
    m = len(rhs)
    rows = slab.split_rows_by_nnz(row_ptr, world, tile)
    halos = slab.halo_ranges(row_ptr, col, rows)
    recv, send = slab.halo_plan(rows, halos)
    ra, rb = rows[rank]
    d = np.array([val[row_ptr[r]:row_ptr[r + 1]][col[row_ptr[r]:row_ptr[r + 1]] == r][0] for r in range(m)])

    def halo(v):
        reqs = []
        for peer, b, e in send[rank]:
            reqs.append(dist.isend(torch.from_numpy(v[b:e].copy()), peer))
        bufs = []
        for peer, b, e in recv[rank]:
            t = torch.empty(e - b, dtype=torch.float64)
            reqs.append(dist.irecv(t, peer)); bufs.append((b, e, t))
        for r_ in reqs:
            r_.wait()
        for b, e, t in bufs:
            v[b:e] = t.numpy()

    import scipy.sparse as ss
    A_own = ss.csr_matrix((val, col, row_ptr), shape=(m, m))[ra:rb]

    def spmv_own(v):
        return A_own @ v          # only stored columns are read: a column outside own rows + halo is NaN

    def allsum(a):
        t = torch.tensor([a], dtype=torch.float64); dist.all_reduce(t); return float(t[0])

    x = np.zeros(m); p = np.full(m, np.nan)          # NaN outside own rows + halo: any stray read is caught
    r = rhs[ra:rb].copy()
    z = r / d[ra:rb]
    p[ra:rb] = z
    rz = allsum(r @ z); bb = allsum(rhs[ra:rb] @ (rhs[ra:rb] / d[ra:rb]))
    it = 0
    while rz > 1e-24 * bb and it < 5000:
        halo(p)
        Ap = spmv_own(p)
        alpha = rz / allsum(p[ra:rb] @ Ap)
        x[ra:rb] += alpha * p[ra:rb]
        r -= alpha * Ap
        z = r / d[ra:rb]
        rzn = allsum(r @ z)
        p[ra:rb] = z + (rzn / rz) * p[ra:rb]
        rz = rzn; it += 1
    # all-gather of the solution in equal padded chunks (the pairwise stage's exchange)
    pieces = []
    for q in range(world):
        t = torch.from_numpy(x[rows[q][0]:rows[q][1]].copy()) if q == rank else torch.zeros(rows[q][1] - rows[q][0], dtype=torch.float64)
        dist.broadcast(t, src=q)
        pieces.append(t)
    sol = np.concatenate([t.numpy() for t in pieces])
    np.save(os.path.join(out_dir, f"sol_{rank}.npy"), sol)
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([it, ra, rb, len(recv[rank]), len(send[rank])]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_plan_drives_distributed_pcg(tmp_path, world):
    import torch.multiprocessing as mp
    from devicekmc_b200 import slab
    tile = 1984
    port = _free_port()
    mp.spawn(_worker, args=(world, port, tile, str(tmp_path)), nprocs=world, join=True)
    row_ptr, col, val, rhs = _system()
    import scipy.sparse as ss
    import scipy.sparse.linalg as sl
    A = ss.csr_matrix((val, col, row_ptr))
    ref = sl.spsolve(A.tocsc(), rhs)
    sols = [np.load(tmp_path / f"sol_{r}.npy") for r in range(world)]
    for s in sols[1:]:
        assert np.array_equal(s, sols[0])                      # every rank holds the same full vector
    assert np.abs(sols[0] - ref).max() / np.abs(ref).max() < 1e-6
    metas = [np.load(tmp_path / f"meta_{r}.npy") for r in range(world)]
    assert len({int(mt[0]) for mt in metas}) == 1              # identical iteration counts (same scalars)
    assert metas[0][1] == 0 and metas[-1][2] == len(rhs)       # rows cover [0, m)
    for a, b in zip(metas[:-1], metas[1:]):
        assert a[2] == b[1]


def test_partition_helpers():
    from devicekmc_b200 import slab
    assert slab.split_even(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.split_even(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert slab.chunk_padded(10, 4) == 3
    # rows by whole tiles: boundaries are tile starts (first row whose row_ptr >= t * tile)
    row_ptr = np.arange(0, 101 * 7, 7)                          # 100 rows of 7 nnz
    rows = slab.split_rows_by_nnz(row_ptr, 4, 64)
    assert rows[0][0] == 0 and rows[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(rows[:-1], rows[1:]))
    for a, b in rows[:-1]:
        t = row_ptr[b] // 64
        assert row_ptr[b] >= t * 64 and row_ptr[b - 1] < t * 64
    # banded matrix: each rank needs only its neighbours; an unordered one needs everybody
    n = 100
    col = np.concatenate([np.clip(np.arange(r - 3, r + 4), 0, n - 1) for r in range(n)])
    halos = slab.halo_ranges(row_ptr, col, rows)
    recv, send = slab.halo_plan(rows, halos)
    for r in range(4):
        assert {q for q, _, _ in recv[r]} == {q for q in (r - 1, r + 1) if 0 <= q < 4}
        for q, b, e in recv[r]:
            assert (r, b, e) in send[q] and rows[q][0] <= b < e <= rows[q][1] and e - b <= 3


def test_owner_of_target_matches_global_scan():
    from devicekmc_b200 import slab
    rng = np.random.default_rng(3)
    rates = np.where(rng.random(1000) < 0.05, rng.integers(1, 100, 1000), 0).astype(np.float64)
    parts = np.array_split(rates, 4)
    sums = np.array([p.sum() for p in parts])
    cum = np.cumsum(rates)
    for u in [0.0, 0.1, 0.5, 0.999, rng.random()]:
        r, local, total = slab.owner_of_target(sums, u)
        gidx = int(np.searchsorted(cum, u * cum[-1], side="right"))     # the reference's upper_bound
        lidx = int(np.searchsorted(np.cumsum(parts[r]), local, side="right"))
        assert total == cum[-1] and sum(len(p) for p in parts[:r]) + lidx == gidx
    assert slab.owner_of_target(np.zeros(4), 0.3)[0] == -1
