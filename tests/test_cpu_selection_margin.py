"""CPU model of the event selection of devicekmc_b200/csrc/events.cu (select_walk): the 32-ary
hierarchy of warp scans, the error bound from the exponent histogram and the margin test that decides
between the fast walk and the exact sequential replay.  The property the kernel relies on:

    whenever the walk does NOT ask for the exact replay, it selects the entry that the reference's
    strictly sequential prefix sum + std::upper_bound (utils.h:91-99, KMCProcess.cpp:306-311) selects.

The model reproduces the kernel's association order bit for bit (Hillis-Steele warp scans, lane-local
row sums) in numpy float64 and is driven with adversarial targets a few ulp on either side of the
sequential prefix sums, where the two summation orders can disagree."""
import numpy as np
import pytest


def warp_scan(v):
    """warp_inclusive_scan (scan.cuh: shfl_up by 1, 2, 4, 8, 16), on the last axis (length 32)"""
    v = np.array(v, dtype=np.float64, copy=True)
    o = 1
    while o < 32:
        t = v.copy()
        v[..., o:] = t[..., o:] + t[..., :-o]
        o <<= 1
    return v


class Model:
    def __init__(self, table):
        self.N, self.nn = table.shape
        self.c = (self.nn + 31) // 32
        self.tab = table
        pad = np.zeros((self.N, 32 * self.c))
        pad[:, :self.nn] = table
        self.lanes = pad.reshape(self.N, 32, self.c)
        # row_scan: lane-local sequential sum of c slots, then the warp scan; row sum = lane 31
        lane_sum = self.lanes[:, :, 0].copy()
        for t in range(1, self.c):
            lane_sum = lane_sum + self.lanes[:, :, t]
        self.row_inc = warp_scan(lane_sum)
        self.levels = [self.row_inc[:, 31].copy()]
        while len(self.levels[-1]) > 1:
            ch = self.levels[-1]
            n = (len(ch) + 31) // 32
            p = np.zeros(n * 32); p[:len(ch)] = ch
            self.levels.append(warp_scan(p.reshape(n, 32))[:, 31].copy())
        # exponent histogram of the non-zero rates (rate_rows_kernel) and the tables of event_loop_kernel
        nz = table[table != 0]
        e = (nz.view(np.int64) >> 52) & 0x7FF
        hist = np.bincount(e, minlength=2048).astype(np.float64)
        ub = np.ldexp(1.0, np.minimum(np.arange(2048) + 1, 2046) - 1023)       # upper edge of bucket b
        self.errA = np.concatenate([[0.0], np.cumsum(hist * ub)[:-1]])          # sum_{b' < b} count * 2^(b'+1)
        self.errB = hist[::-1].cumsum()[::-1]                                   # sum_{b' >= b} count
        self.cum_seq = np.cumsum(table.ravel())                                 # the reference: sequential

    def delta(self, psum, factor_E, factor_tau):
        tau = psum * 2.220446049250313e-16 * 1.000001
        bt = (np.float64(tau).view(np.int64) >> 52) & 0x7FF
        return factor_E * (self.errA[bt] + tau * self.errB[bt]) + factor_tau * tau

    def select_walk(self, u, factor_E=8.0, factor_tau=128.0):
        """returns (idx, flag): flag 0 ok, 1 none, 2 needs the exact replay"""
        prefix, g = 0.0, 0
        top = len(self.levels) - 1
        for lev in range(top, -1, -1):
            L = self.levels[lev]
            v = np.zeros(32); seg = L[g * 32:(g + 1) * 32]; v[:len(seg)] = seg
            inc = warp_scan(v)
            if lev == top:
                psum = inc[31]
                number = u * psum
                d = self.delta(psum, factor_E, factor_tau)
            glob = prefix + inc
            hit = np.nonzero(glob > number)[0]
            if len(hit) == 0:
                return -1, (1 if lev == top else 2)
            l = hit[0]
            if l > 0:
                prefix = glob[l - 1]
            g = g * 32 + l
        row = g
        inc = self.row_inc[row]
        excl = np.concatenate([[0.0], inc[:-1]])
        for lane in range(32):
            part = 0.0
            for t in range(self.c):
                prev = part
                part = self.lanes[row, lane, 0] if t == 0 else part + self.lanes[row, lane, t]
                cum = prefix + (excl[lane] + part)
                if cum > number:
                    cb = prefix + (excl[lane] + (0.0 if t == 0 else prev))
                    idx = row * self.nn + lane * self.c + t
                    ok = (number - cb > d) and (cum - number > d)
                    return idx, (0 if ok else 2)
        return -1, 2

    def select_seq(self, u):
        number = u * self.cum_seq[-1]
        k = int(np.searchsorted(self.cum_seq, number, side="right"))
        return k if k < len(self.cum_seq) else -1


def make_table(rng, N, nn, kind):
    t = np.zeros((N, nn))
    if kind == "kmc":          # a third of the rows carry tiny rates, a few rows carry rates that matter
        rows = rng.random(N) < 0.33
        t[rows] = np.where(rng.random((rows.sum(), nn)) < 0.15, np.exp(rng.uniform(-160, -60, (rows.sum(), nn))), 0.0)
        big = rng.choice(N, max(4, N // 40), replace=False)
        t[big] = np.where(rng.random((len(big), nn)) < 0.2, np.exp(rng.uniform(-12, 25, (len(big), nn))), 0.0)
    elif kind == "flat":       # thousands of entries of similar size: the worst case for rounding drift
        t = np.where(rng.random((N, nn)) < 0.5, rng.uniform(0.5, 2.0, (N, nn)), 0.0)
    elif kind == "wide":       # every magnitude at once
        t = np.where(rng.random((N, nn)) < 0.3, np.exp(rng.uniform(-300, 300, (N, nn))), 0.0)
    return t


@pytest.mark.parametrize("kind", ["kmc", "flat", "wide"])
def test_fast_walk_agrees_with_sequential_search_whenever_it_does_not_fall_back(kind):
    rng = np.random.default_rng({"kmc": 1, "flat": 2, "wide": 3}[kind])
    N, nn = 2048, 51
    m = Model(make_table(rng, N, nn, kind))
    psum = m.cum_seq[-1]
    nzpos = np.nonzero(m.tab.ravel())[0]
    fast = fallback = 0
    # adversarial targets: a few ulp around the sequential prefix sums, plus uniform ones
    us = list(rng.random(300))
    for k in rng.choice(nzpos, 500):
        for s in (-64, -3, -1, 0, 1, 3, 64):
            b = m.cum_seq[k]
            us.append(min(max((b + s * np.spacing(b)) / psum, 0.0), np.nextafter(1.0, 0.0)))
    for u in us:
        idx, flag = m.select_walk(u)
        if flag == 0:
            fast += 1
            assert idx == m.select_seq(u), (kind, u)
        elif flag == 2:
            fallback += 1
    assert fast > 200                              # uniform targets take the fast walk
    assert fallback > 0                            # targets within a few ulp of a boundary must not


def test_margin_is_tighter_than_the_first_version_but_not_vacuous():
    """the same table: delta = 8 E' + 128 tau replays several times less often than 64 (E' + tau)"""
    rng = np.random.default_rng(7)
    m = Model(make_table(rng, 2048, 51, "flat"))
    psum = m.levels[-1][0]
    d_new, d_old = m.delta(psum, 8.0, 128.0), m.delta(psum, 64.0, 64.0)
    assert 4.0 < d_old / d_new <= 8.0
    # and the bound still dominates the observed difference between the two summation orders
    prefixes_walk = []
    for u in rng.random(200):
        idx, flag = m.select_walk(u, 0.0, 0.0)     # no margin: compare the raw decisions
        prefixes_walk.append(idx == m.select_seq(u))
    assert np.mean(prefixes_walk) > 0.95


def test_absorbed_entries_can_be_skipped_in_the_exact_replay():
    """Round-2 groundwork for select_exact: the sequential sum never decreases, so an
    entry smaller than a quarter ulp of ANY earlier prefix (here: the running sum at the start of its
    batch) is absorbed — acc + p == acc — and a parallel pre-pass may drop it without changing a single
    bit of the sequential sums.  Checked on tables with 300 decades of dynamic range."""
    for seed, kind in ((1, "kmc"), (3, "wide"), (5, "kmc")):
        rng = np.random.default_rng(seed)
        v = make_table(rng, 2048, 51, kind).ravel()
        v = v[v != 0]
        full = np.cumsum(v)
        acc, kept, out = 0.0, 0, np.empty_like(v)
        B = 512                                            # entries per batch
        for b0 in range(0, len(v), B):
            lb = acc                                       # lower bound of every prefix inside the batch
            thr = np.spacing(lb) / 4 if lb > 0 else 0.0
            for k in range(b0, min(len(v), b0 + B)):
                if v[k] >= thr:                            # survives the pre-pass: the one adder adds it
                    acc = acc + v[k]
                    kept += 1
                out[k] = acc
        assert np.array_equal(out, full)
        if kind == "kmc":
            assert kept < 0.5 * len(v)                     # most of the tiny generation rates drop out
