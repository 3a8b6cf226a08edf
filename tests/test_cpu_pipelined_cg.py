"""CPU model of the recurrences of devicekmc_b200/csrc/pcg_pipelined.cuh (no GPU): pipelined CG with
u = M^-1 r and q = M^-1 s applied on the fly and the cluster sums W^T z, W^T s, W^T r, W^T w carried by their own
recurrences, M^-1 = D^-1 + W E^-1 W^T.  The model must (a) solve the system, (b) take the same iterates as the textbook
preconditioned CG of iterative_solvers_gpu.cu:424-455 in exact arithmetic (here: to rounding), and (c) show the
property the host logic relies on: its TRUE residual stagnates above the classic recurrence's, so it is used for the
restarts' correction solves only."""
import numpy as np
import scipy.sparse as sp


def _system(n=400, n_clusters=6, seed=0):
    rng = np.random.default_rng(seed)
    # graph Laplacian of a ring with random chords, conductances 1e-8 except inside a few "clusters" (1.0),
    # grounded at a handful of nodes: the structure of K (potential_solver.cpp:325-372)
    rows, cols, vals = [], [], []
    edges = {(i, (i + 1) % n) for i in range(n)} | {tuple(sorted(rng.choice(n, 2, replace=False))) for _ in range(3 * n)}
    members = [np.arange(40 * c + 5, 40 * c + 5 + 3) for c in range(n_clusters)]
    strong = {(int(a), int(b)) for m in members for a in m for b in m if a < b}
    edges |= strong
    for a, b in edges:
        g = 1.0 if (min(a, b), max(a, b)) in strong else 1e-8
        rows += [a, b, a, b]; cols += [b, a, a, b]; vals += [-g, -g, g, g]
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n)).tolil()
    for k in rng.choice(n, 12, replace=False):
        A[k, k] += 1.0                                    # contact conductances on the diagonal
    A = A.tocsr()
    W = sp.csr_matrix((np.ones(sum(len(m) for m in members)),
                       (np.concatenate(members), np.repeat(np.arange(n_clusters), [len(m) for m in members]))), shape=(n, n_clusters))
    b = rng.standard_normal(n) * (A.diagonal() > 0.5)     # right-hand side on the contact rows, like -K_sub V
    return A, W, b


def _precond(A, W):
    dinv = 1.0 / A.diagonal()
    E = np.asarray((W.T @ A @ W).todense()).diagonal()    # clusters do not touch each other
    return dinv, 1.0 / E


def pipelined(A, W, b, tol, max_iter=4000):
    """the kernel's loop: vectors x, r, w, p, s, z + gather vector m; cluster recurrences cz, cs, cr, cw"""
    dinv, wE = _precond(A, W)
    n = len(b)
    x = np.zeros(n); r = b - A @ x
    cr = W.T @ r
    u = dinv * r + W @ (wE * cr)
    w = A @ u
    cw = W.T @ w
    gamma, delta, bb = r @ u, w @ u, b @ (dinv * b + W @ (wE * (W.T @ b)))
    m = dinv * w + W @ (wE * cw)
    nvec = A @ m
    cn = W.T @ nvec
    z = np.zeros(n); s = np.zeros(n); p = np.zeros(n)
    cz = np.zeros(W.shape[1]); cs = np.zeros(W.shape[1])
    alpha = beta = 0.0
    gamma_old = 1.0
    for it in range(max_iter):
        if gamma <= tol * tol * bb:
            break
        if it == 0:
            beta, alpha = 0.0, gamma / delta
        else:
            beta = gamma / gamma_old
            alpha = gamma / (delta - beta * gamma / alpha)
        cz = cn + beta * cz; cs = cw + beta * cs
        cr_old = cr.copy(); cr = cr - alpha * cs; cw = cw - alpha * cz
        u_old = dinv * r + W @ (wE * cr_old)
        z = nvec + beta * z; s = w + beta * s; p = u_old + beta * p
        x = x + alpha * p; r = r - alpha * s; w = w - alpha * z
        u = dinv * r + W @ (wE * cr)
        gamma_old = gamma
        gamma, delta = r @ u, w @ u
        m = dinv * w + W @ (wE * cw)
        nvec = A @ m
        cn = W.T @ nvec
    pipelined.last = dict(r=r, w=w, cr=cr, cw=cw, u=u)
    return x, it


def classic(A, W, b, tol, max_iter=4000):
    dinv, wE = _precond(A, W)
    Minv = lambda v: dinv * v + W @ (wE * (W.T @ v))
    x = np.zeros(len(b)); r = b.copy(); zv = Minv(r); p = zv.copy(); rz = r @ zv; bb = b @ Minv(b)
    for it in range(max_iter):
        if rz <= tol * tol * bb:
            return x, it
        Ap = A @ p; a = rz / (p @ Ap); x = x + a * p; r = r - a * Ap; zv = Minv(r); rzn = r @ zv; p = zv + (rzn / rz) * p; rz = rzn
    return x, max_iter


def test_pipelined_recurrences_solve_the_system_like_the_classic_cg():
    A, W, b = _system()
    xs = np.linalg.solve(A.toarray(), b)
    xp, itp = pipelined(A, W, b, 1e-8)
    xc, itc = classic(A, W, b, 1e-8)
    assert abs(itp - itc) <= 3                                  # the same iteration, rearranged
    scale = np.abs(xs).max()
    assert np.abs(xp - xs).max() <= 1e-5 * scale and np.abs(xc - xs).max() <= 1e-5 * scale
    assert np.abs(xp - xc).max() <= 1e-6 * scale


def test_cluster_recurrences_track_the_true_cluster_sums():
    """the carried W^T r and W^T w equal W^T of the recurrence vectors r and w (so u = D^-1 r + W E^-1 (W^T r) applied
    on the fly IS M^-1 r), and w tracks A u"""
    A, W, b = _system(seed=3)
    x, it = pipelined(A, W, b, 1e-6)
    L = pipelined.last
    assert it > 10
    assert np.abs(W.T @ L["r"] - L["cr"]).max() <= 1e-10 * (np.abs(L["cr"]).max() + np.abs(W.T @ b).max() + 1e-300) + 1e-18
    assert np.abs(W.T @ L["w"] - L["cw"]).max() <= 1e-8 * (np.abs(L["cw"]).max() + 1e-300) + 1e-18
    assert np.abs(A @ L["u"] - L["w"]).max() <= 1e-4 * np.abs(L["w"]).max() + 1e-14   # w drifts from A u: the pipelined floor


def test_pipelined_true_residual_stagnates_above_the_classic_one():
    A, W, b = _system(n=800, n_clusters=10, seed=1)
    dinv, wE = _precond(A, W)
    Minv = lambda v: dinv * v + W @ (wE * (W.T @ v))
    rel = lambda x: np.sqrt(((b - A @ x) @ Minv(b - A @ x)) / (b @ Minv(b)))
    xp, itp = pipelined(A, W, b, 1e-15, max_iter=200)
    xc, itc = classic(A, W, b, 1e-15, max_iter=200)
    assert itc < 200 and rel(xc) < 1e-7                         # the classic recurrence gets there ...
    assert itp >= 199                                           # ... the pipelined one never meets a 1e-15 tolerance,
    assert rel(xp) < 1e-4                                       # although it has converged to its floor,
    assert rel(xp) > 10 * rel(xc)                               # which lies above the classic one's: restarts only
