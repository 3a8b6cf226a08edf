"""Host logic of the bias-point loop (slab.bias_loop = kmc_main.cpp:136-279) on a stand-in simulation object (no GPU):
syncs per bias point, KMC time accumulated until t_switch, the bounds on the work, and the rescaled warm start."""
import numpy as np

from devicekmc_b200 import slab


class _Vec:
    def __init__(self):
        self.scale = 1.0

    def mul_(self, f):
        self.scale *= f


class _Buf:
    def __init__(self, log):
        self.log = log
        self.site_potential_boundary = _Vec()

    def sync_HostToGPU(self, dev):
        self.log.append("h2d")

    def sync_GPUToHost(self, dev):
        self.log.append("d2h")


class _Sim:
    """every step advances the KMC time by dt(Vd): long residence times at low bias, short ones at high bias"""

    def __init__(self):
        self.log = []
        self.buf = _Buf(self.log)
        self.dev = None

    def step(self, Vd, record_events=0):
        self.log.append(("step", Vd))
        return {"step_time": 2e-3 if Vd < 1.0 else 4e-4, "events": 1}


def test_one_step_per_point_at_low_bias_and_several_at_high_bias():
    s = _Sim()
    st = slab.bias_loop(s, [0.0, 0.5, 1.5], 1e-3, scale_warm_start=False)
    assert [(t["Vd"], t["bias_point"]) for t in st] == [(0.0, 0), (0.5, 1), (1.5, 2), (1.5, 2), (1.5, 2)]
    # kmc_main.cpp:172 / :282: one upload before and one download after the steps of a bias point
    assert s.log == ["h2d", ("step", 0.0), "d2h", "h2d", ("step", 0.5), "d2h", "h2d", ("step", 1.5), ("step", 1.5), ("step", 1.5), "d2h"]


def test_bounds_on_the_work():
    s = _Sim()
    st = slab.bias_loop(s, [2.0, 3.0], 1e-3, max_steps_per_point=2, scale_warm_start=False)
    assert [t["bias_point"] for t in st] == [0, 0, 1, 1]
    s = _Sim()
    st = slab.bias_loop(s, [2.0, 3.0], 1e-3, max_steps=3, scale_warm_start=False)
    assert [t["bias_point"] for t in st] == [0, 0, 0]
    assert s.log[-1] == "d2h"                       # the state is back on the host when the loop stops early


def test_warm_start_is_rescaled_by_the_voltage_ratio_only_when_it_can_be():
    s = _Sim()
    slab.bias_loop(s, [0.0, 0.5, 1.0, 1.0 + 1e-12, 2.0], [1e-3] * 5, max_steps_per_point=1)
    # 0 -> 0.5: nothing to scale from; 0.5 -> 1.0 -> 1.0+1e-12 -> 2.0: the product of the ratios
    assert np.isclose(s.buf.site_potential_boundary.scale, 2.0 / 0.5)


def test_iv_ramp_is_the_set_reset_sweep():
    r = slab.iv_ramp(200, 4.0)
    assert len(r) == 200 and r[0] == 0.0 and r[-1] == 0.0 and r.max() == 4.0
    assert np.all(np.diff(r[:101]) > 0) and np.all(np.diff(r[100:]) < 0)
    assert sorted(slab.RAMP_TILES) == [1, 2, 4, 8] and slab.RAMP_TILES[8] == (21, 20)     # 8 GPUs: the 4 M-site device
