"""Host-side mirror of the reference interface for the field-and-rate hot path.

Same names and argument meaning as the reference's C++ host classes, so the parity tests read
like the reference's own flow (kmc_main.cpp:175-279):

    Device(xyz_files, p)              Device.cpp:17        site arrays + neighbour graph
    Device.makeSubstoichiometric      Device.cpp:202
    Device.updateCharge(gpubuf, ..)   potential_solver.cpp:142
    Device.updatePotential(..)        potential_solver.cpp:232
    Device.setLaplacePotential(..)    potential_solver.cpp:4       CB edge, once per bias point
    KMCProcess(device, freq)          KMCProcess.cpp:17    layers, site->layer, KMC RNG
    KMCProcess.executeKMCStep(..)     KMCProcess.cpp:259
    GPUBuffers(...)                   gpu_buffers.h:73     device-resident mirror of Device
    GPUBuffers.sync_HostToGPU / sync_GPUToHost             gpu_buffers.cpp:10-55
    RandomNumberGenerator             random_num.h:4-23

PyTorch is used only for device memory and streams; all arithmetic of the path happens in
libdkmc_b200.so (hand-written sm_100a kernels) behind the C-ABI of include/dkmc.h.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import Sparsity, SolveInfo, SolverOpts, StepInfo, check

# ELEMENT / EVENTTYPE (utils.h:37-60)
DEFECT, OXYGEN_DEFECT, VACANCY, O_EL, Hf_EL, Ni_EL, Ti_EL, Pt_EL, N_EL, NULL_ELEMENT = range(10)
VACANCY_GENERATION, VACANCY_RECOMBINATION, VACANCY_DIFFUSION, ION_DIFFUSION, NULL_EVENT = range(5)

_ELEMENT_OF = {"d": DEFECT, "Od": OXYGEN_DEFECT, "V": VACANCY, "O": O_EL, "Hf": Hf_EL, "Ni": Ni_EL,
               "Ti": Ti_EL, "Pt": Pt_EL, "N": N_EL}
_NAME_OF = {v: k for k, v in _ELEMENT_OF.items()}


def update_element(name: str) -> int:
    """utils.cpp:5-29"""
    try:
        return _ELEMENT_OF[name]
    except KeyError:
        raise ValueError(f"Unknown element type in update_element!: {name}")


def return_element(e: int) -> str:
    return _NAME_OF.get(int(e), "")


def read_xyz(filename: str):
    """utils.cpp:72-98: line 1 = N, line 2 = comment, then `element x y z` rows."""
    with open(filename) as f:
        n = int(f.readline().split()[0])
        f.readline()
        el = np.empty(n, np.int32)
        xyz = np.empty((n, 3), np.float64)
        for i in range(n):
            t = f.readline().split()
            el[i] = update_element(t[0])
            xyz[i] = (float(t[1]), float(t[2]), float(t[3]))
    return el, xyz[:, 0].copy(), xyz[:, 1].copy(), xyz[:, 2].copy()


def write_xyz(filename: str, element, x, y, z, extra=None):
    """Device::writeSnapshot layout (Device.cpp:236-252)"""
    with open(filename, "w") as f:
        f.write(f"{len(x)}\n\n")
        for i in range(len(x)):
            tail = "" if extra is None else "   " + "   ".join(repr(float(c[i])) for c in extra)
            f.write(f"{return_element(element[i])}   {float(x[i])!r}   {float(y[i])!r}   {float(z[i])!r}{tail}\n")


def write_snapshot(path: str, element, x, y, z, potential, power=None):
    """The file Device::writeSnapshot produces (Device.cpp:236-252): `element x y z potential power`,
    three blanks between columns, default ostream formatting (= %g, 6 significant digits)."""
    names = [return_element(e) for e in range(NULL_ELEMENT + 1)]
    n = len(x)
    pw = np.zeros(n) if power is None else power
    with open(path, "w") as f:
        f.write(f"{n}\n\n")
        for lo in range(0, n, 65536):
            hi = min(n, lo + 65536)
            f.write("".join(
                f"{names[e]}   {a:g}   {b:g}   {c:g}   {v:g}   {w:g}\n"
                for e, a, b, c, v, w in zip(element[lo:hi].tolist(), x[lo:hi].tolist(), y[lo:hi].tolist(),
                                            z[lo:hi].tolist(), potential[lo:hi].tolist(), pw[lo:hi].tolist())))


class RandomNumberGenerator:
    """std::mt19937 + std::uniform_real_distribution<double>(0,1) as libstdc++ implements it
    (generate_canonical<double,53>: two 32-bit draws per double).  random_num.h:4-23."""

    def __init__(self, seed: int = 0):
        self.setSeed(seed)

    def setSeed(self, seed: int):
        self._bg = np.random.MT19937()
        self._bg._legacy_seeding(int(seed) & 0xFFFFFFFF)  # init_genrand(seed) == std::mt19937(seed)

    def _canonical(self, raw: np.ndarray) -> np.ndarray:
        lo = raw[0::2].astype(np.float64)
        hi = raw[1::2].astype(np.float64)
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        r[r >= 1.0] = np.nextafter(1.0, 0.0)
        return r

    def getRandomNumber(self) -> float:
        return float(self._canonical(self._bg.random_raw(2))[0])

    def getRandomNumbers(self, n: int) -> np.ndarray:
        return self._canonical(self._bg.random_raw(2 * n))

    def peek(self, n: int) -> np.ndarray:
        """the next n numbers without consuming them"""
        st = self._bg.state
        out = self.getRandomNumbers(n)
        self._bg.state = st
        return out

    def advance(self, n: int):
        if n > 0:
            self._bg.random_raw(2 * n)


@dataclasses.dataclass
class Layer:
    """utils.h:63-72 / structure_input.h:8-50"""
    type: str
    E_gen_0: float
    E_rec_1: float
    E_diff_2: float
    E_diff_3: float
    start_x: float
    end_x: float


# structure_input.h:12-50 (compile-time constants of the reference)
DEFAULT_LAYERS = [
    Layer("contact", 0.0, 0.0, 0.0, 0.76, -22.0, 0.0),
    Layer("interface", 3.93, 0.0, 1.09, 0.76, 0.0, 3.0),
    Layer("oxide", 3.93, 0.0, 1.09, 0.76, 3.0, 48.1431),
    Layer("interface", 1.66, 0.0, 1.09, 0.76, 48.1431, 52.6431),
    Layer("contact", 1.73, 0.0, 0.0, 2.8, 52.6431, 90.0),
]
RND_SEED_KMC = 1  # structure_input.h:8


@dataclasses.dataclass
class KMCParameters:
    """The subset of input_parser.h's KMCParameters that feeds the hot path."""
    lattice: Sequence[float] = (108.97557, 25.575, 25.575)
    pbc: int = 0
    nn_dist: float = 3.5
    sigma: float = 3.5e-10
    epsilon: float = 23.0
    background_temp: float = 300.0
    freq: float = 1e14
    metals: Sequence[int] = (Ti_EL, N_EL)
    num_atoms_first_layer: int = 144
    num_atoms_contact: int = 144
    rnd_seed: int = 4
    pristine: int = 1
    initial_vacancy_concentration: float = 0.05
    high_G: float = 1.0    # input_parser.cpp:392
    low_G: float = 1e-8    # input_parser.cpp:393
    q: float = 1.60217663e-19   # input_parser.h: elementary charge [C]
    V_switch: Sequence[float] = (0.0,)
    t_switch: Sequence[float] = (1e-3,)
    restart_xyz_file: str = ""

    @property
    def k(self) -> float:
        return 8.987552e9 / self.epsilon  # Device.cpp:35

    @staticmethod
    def from_file(path: str) -> "KMCParameters":
        """`key = value // comment` files of the reference (input_parser.cpp:3-249); only the keys
        the hot path consumes are read."""
        kv: Dict[str, str] = {}
        with open(path) as f:
            for line in f:
                line = line.split("//")[0].strip()
                if "=" not in line:
                    continue
                k_, v_ = line.split("=", 1)
                kv.setdefault(k_.strip(), v_.strip())

        def nums(key):
            return [float(t) for t in re.split(r"[,\s]+", kv[key]) if t]

        p = KMCParameters()
        if "lattice" in kv: p.lattice = tuple(nums("lattice"))
        if "pbc" in kv: p.pbc = int("1" in kv["pbc"])
        if "nn_dist" in kv: p.nn_dist = nums("nn_dist")[-1]
        if "sigma" in kv: p.sigma = nums("sigma")[-1]
        if "epsilon" in kv: p.epsilon = nums("epsilon")[-1]
        if "background_temp" in kv: p.background_temp = nums("background_temp")[-1]
        if "attempt_frequency" in kv: p.freq = nums("attempt_frequency")[-1]
        if "metals" in kv: p.metals = tuple(update_element(t) for t in kv["metals"].split())
        if "num_atoms_first_layer" in kv: p.num_atoms_first_layer = int(nums("num_atoms_first_layer")[-1])
        if "num_atoms_contact" in kv: p.num_atoms_contact = int(nums("num_atoms_contact")[-1])
        if "rnd_seed" in kv: p.rnd_seed = int(nums("rnd_seed")[-1])
        if "pristine" in kv: p.pristine = int("1" in kv["pristine"])
        if "initial_vacancy_concentration" in kv:
            p.initial_vacancy_concentration = nums("initial_vacancy_concentration")[-1]
        if "V_switch" in kv: p.V_switch = tuple(nums("V_switch"))
        if "t_switch" in kv: p.t_switch = tuple(nums("t_switch"))
        if "restart_xyz_file" in kv: p.restart_xyz_file = kv["restart_xyz_file"]
        return p


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("devicekmc-b200 needs a CUDA device: the hot path has no CPU fallback")
    return torch


class Context:
    """dkmc_ctx: workspace arena + stream of the C-ABI library."""

    def __init__(self):
        self.lib = _capi.load()
        h = C.c_void_p()
        check(self.lib.dkmc_ctx_create(C.byref(h)))
        self.h = h
        self.use_current_stream()

    def use_current_stream(self):
        torch = _torch()
        check(self.lib.dkmc_ctx_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def set_pairwise_incremental(self, refresh_every: int):
        """opt-in: phi_c is updated by the charge differences since the previous step,
        with a full sum every `refresh_every` steps; 0 switches it off"""
        check(self.lib.dkmc_ctx_set_pairwise_incremental(self.h, int(refresh_every)))

    def pairwise_incremental_counts(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        check(self.lib.dkmc_pairwise_incremental_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self) -> int:
        n = C.c_longlong(0)
        check(self.lib.dkmc_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.dkmc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


class Device:
    """Host view of the device structure (Device.h).  The neighbour graph is built on the GPU by
    the cell-list kernel instead of the reference's O(N^2) loop (Device.cpp:98-136)."""

    def __init__(self, xyz_files, p: KMCParameters, ctx: Optional[Context] = None, arrays=None):
        torch = _torch()
        self.ctx = ctx or Context()
        self.random_generator = RandomNumberGenerator(p.rnd_seed)
        if arrays is not None:
            el, x, y, z = arrays
        else:
            parts = [read_xyz(f) for f in xyz_files]
            el = np.concatenate([q[0] for q in parts]); x = np.concatenate([q[1] for q in parts])
            y = np.concatenate([q[2] for q in parts]); z = np.concatenate([q[3] for q in parts])
        self.N = int(len(x))
        self.site_element = np.ascontiguousarray(el, np.int32)
        self.site_x, self.site_y, self.site_z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
        self.lattice = np.asarray(p.lattice, np.float64)
        self.pbc = int(p.pbc)
        self.nn_dist = float(p.nn_dist)
        self.sigma = float(p.sigma)
        self.k = p.k
        self.T_bg = float(p.background_temp)
        # neighbour graph on the device
        dev = torch.device("cuda")
        self._dx = torch.from_numpy(self.site_x).to(dev)
        self._dy = torch.from_numpy(self.site_y).to(dev)
        self._dz = torch.from_numpy(self.site_z).to(dev)
        lib = self.ctx.lib
        lat = self.lattice.ctypes.data_as(C.c_void_p)
        nn = C.c_int(0)
        check(lib.dkmc_neighbor_count(self.ctx.h, self.N, _ptr(self._dx), _ptr(self._dy), _ptr(self._dz), lat,
                                      self.pbc, self.nn_dist, C.byref(nn)))
        self.max_num_neighbors = nn.value
        self._d_neigh = torch.empty(self.N * nn.value, dtype=torch.int32, device=dev)
        check(lib.dkmc_neighbor_fill(self.ctx.h, self.N, _ptr(self._dx), _ptr(self._dy), _ptr(self._dz), lat,
                                     self.pbc, self.nn_dist, nn.value, _ptr(self._d_neigh)))
        self.neigh_idx = self._d_neigh.cpu().numpy()
        self.site_charge = np.zeros(self.N, np.int32)
        self.site_potential_boundary = np.zeros(self.N, np.float64)
        self.site_potential_charge = np.zeros(self.N, np.float64)
        self.site_temperature = np.full(self.N, self.T_bg, np.float64)
        self.site_CB_edge = np.zeros(self.N, np.float64)
        self.updateAtomLists()

    def updateAtomLists(self):
        """Device.cpp:138-172: atoms = sites that are neither `d` nor `Od`"""
        self.atom_ind = np.nonzero((self.site_element != DEFECT) & (self.site_element != OXYGEN_DEFECT))[0]
        self.N_atom = int(len(self.atom_ind))

    def makeSubstoichiometric(self, vacancy_concentration: float):
        """Device.cpp:202-233, same random stream (device RNG, seed p.rnd_seed)"""
        num_O = int(np.count_nonzero(self.site_element == O_EL))
        num_V_add = int(vacancy_concentration * num_O)
        atom_element = self.site_element[self.atom_ind].copy()
        while num_V_add > 0:
            loc = int(self.random_generator.getRandomNumber() * self.N_atom)
            if atom_element[loc] == O_EL:
                atom_element[loc] = VACANCY
                self.site_element[self.atom_ind[loc]] = VACANCY
                num_V_add -= 1

    # ---- the path (dispatch to the GPU entry points, as the reference's #ifdef USE_CUDA branches)
    def updateCharge(self, gpubuf: "GPUBuffers", metals=None) -> dict:
        lib = self.ctx.lib
        check(lib.dkmc_update_charge(self.ctx.h, _ptr(gpubuf.site_element), _ptr(gpubuf.site_charge),
                                     _ptr(gpubuf.neigh_idx), gpubuf.N_, gpubuf.nn_, _ptr(gpubuf.metal_types),
                                     gpubuf.num_metal_types_))
        return {}

    def updatePotential(self, gpubuf: "GPUBuffers", p: KMCParameters, Vd: float, kmc_step_count: int = 0,
                        opts: Optional[SolverOpts] = None, n_contact: Optional[int] = None,
                        overlap: bool = True) -> dict:
        """potential_solver.cpp:232-285.  n_contact: contact size; the reference's GPU branch uses
        num_atoms_first_layer (:240-241), its CPU branch num_atoms_contact (:271)."""
        lib = self.ctx.lib
        nc = p.num_atoms_first_layer if n_contact is None else n_contact
        sp = gpubuf.sparsity(nc, nc)
        info = SolveInfo()
        pw_args = (self.ctx.h, self.pbc, gpubuf.N_, _ptr(gpubuf.lattice), _ptr(gpubuf.sigma), _ptr(gpubuf.k),
                   _ptr(gpubuf.site_x), _ptr(gpubuf.site_y), _ptr(gpubuf.site_z), _ptr(gpubuf.site_charge))
        if overlap:
            # the pairwise sum (FP64 pipe) goes to the side stream and shares the SMs with the CG (HBM)
            check(lib.dkmc_poisson_gridless_begin(*pw_args, 0, gpubuf.N_, _ptr(gpubuf.site_potential_charge)))
        st = lib.dkmc_background_potential_sparse(
            self.ctx.h, C.byref(sp), self.N, gpubuf.nn_, _ptr(gpubuf.neigh_idx), nc, nc, float(Vd),
            float(p.high_G), float(p.low_G), _ptr(gpubuf.site_element), _ptr(gpubuf.site_charge),
            _ptr(gpubuf.metal_types), gpubuf.num_metal_types_, _ptr(gpubuf.site_potential_boundary),
            C.byref(opts) if opts is not None else None, C.byref(info))
        pw_ms = C.c_double(0.0)
        if overlap:
            check(lib.dkmc_poisson_gridless_join(self.ctx.h, C.byref(pw_ms)))
        check(st, allow=(_capi.DKMC_ERR_NOT_CONVERGED,))
        if not overlap:
            torch = _torch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.dkmc_poisson_gridless(*pw_args, _ptr(gpubuf.site_potential_charge)))
            e1.record()
            e1.synchronize()
            pw_ms.value = e0.elapsed_time(e1)
        return {"pairwise_ms": pw_ms.value, "cg_iterations": info.iterations, "cg_rel_residual": info.rel_residual,
                "cg_est_error": info.est_error, "cg_refinements": info.refinements,
                "cg_converged": st == _capi.DKMC_OK, "assemble_ms": info.assemble_ms, "solve_ms": info.solve_ms,
                "overlap": bool(overlap)}


    def writeSnapshot(self, filename: str, foldername: str):
        """Device.cpp:236-252, from the host arrays (call sync_GPUToHost first, as kmc_main.cpp:196-201
        does, or use GPUBuffers.snapshot_begin for a copy that does not stall the step)"""
        import os
        write_snapshot(os.path.join(".", foldername, filename), self.site_element, self.site_x, self.site_y, self.site_z,
                       self.site_potential_boundary + self.site_potential_charge, getattr(self, "site_power", None))

    def _set_laplace(self, gpubuf, p, Vd, opts=None):
        lib = self.ctx.lib
        nc = p.num_atoms_first_layer                      # potential_solver.cpp:7-8
        sp = gpubuf.sparsity(nc, nc)
        info = SolveInfo()
        st = lib.dkmc_update_CB_edge_sparse(
            self.ctx.h, C.byref(sp), self.N, nc, nc, float(Vd), float(p.q), float(p.high_G), float(p.low_G),
            _ptr(gpubuf.site_element), _ptr(gpubuf.metal_types), gpubuf.num_metal_types_, _ptr(gpubuf.site_CB_edge),
            C.byref(opts) if opts is not None else None, C.byref(info))
        check(st, allow=(_capi.DKMC_ERR_NOT_CONVERGED,))
        return {"cg_iterations": info.iterations, "cg_est_error": info.est_error, "cg_refinements": info.refinements,
                "cg_converged": st == _capi.DKMC_OK, "assemble_ms": info.assemble_ms, "solve_ms": info.solve_ms}

    def setLaplacePotential(self, gpubuf: "GPUBuffers", p: KMCParameters, Vd: float,
                            opts: Optional[SolverOpts] = None) -> dict:
        """potential_solver.cpp:4-139, GPU branch (:10-19): host -> device sync, the CB-edge solve
        (update_CB_edge_gpu_sparse), device -> host sync.  Called once per bias point
        (kmc_main.cpp:160)."""
        gpubuf.sync_HostToGPU(self, also=("site_CB_edge",))
        out = self._set_laplace(gpubuf, p, Vd, opts)
        gpubuf.sync_GPUToHost(self, also=("site_CB_edge",))
        return out


class GPUBuffers:
    """Device-resident mirror of Device (gpu_buffers.h:12-162): one array per site attribute, the
    1-element scalars the reference keeps on the device, and the CSR index buffers of K."""

    def __init__(self, layers: List[Layer], site_layer, freq: float, device: Device, metals: Sequence[int]):
        torch = _torch()
        dev = torch.device("cuda")
        self.ctx = device.ctx
        self.N_ = device.N
        self.nn_ = device.max_num_neighbors
        self.N_atom_ = device.N_atom
        self.num_metal_types_ = len(metals)
        f64 = dict(dtype=torch.float64, device=dev)
        self.site_x, self.site_y, self.site_z = device._dx, device._dy, device._dz
        self.neigh_idx = device._d_neigh
        self.site_layer = torch.from_numpy(np.ascontiguousarray(site_layer, np.int32)).to(dev)
        self.site_element = torch.empty(self.N_, dtype=torch.int32, device=dev)
        self.site_charge = torch.zeros(self.N_, dtype=torch.int32, device=dev)
        self.site_potential_boundary = torch.zeros(self.N_, **f64)
        self.site_potential_charge = torch.zeros(self.N_, **f64)
        self.site_temperature = torch.full((self.N_,), device.T_bg, **f64)
        self.site_CB_edge = torch.zeros(self.N_, **f64)
        self.metal_types = torch.tensor(list(metals), dtype=torch.int32, device=dev)
        self.sigma = torch.tensor([device.sigma], **f64)
        self.k = torch.tensor([device.k], **f64)
        self.T_bg = torch.tensor([device.T_bg], **f64)
        self.freq = torch.tensor([freq], **f64)
        self.lattice = torch.tensor(list(device.lattice), **f64)
        self.E_host = np.array([[l.E_gen_0 for l in layers], [l.E_rec_1 for l in layers],
                                [l.E_diff_2 for l in layers], [l.E_diff_3 for l in layers]], np.float64)
        # copytoConstMemory (gpu_buffers.h:102)
        e = [np.ascontiguousarray(r) for r in self.E_host]
        check(self.ctx.lib.dkmc_set_layer_energies(self.ctx.h, len(layers), *[a.ctypes.data_as(C.c_void_p) for a in e]))
        self._sparsity: Dict[tuple, Sparsity] = {}
        self._host_xyz = (device.site_x, device.site_y, device.site_z)
        # "auto": inputs whose interior sites are not already in x-major grid-cell order (the reference's order puts
        # lattice atoms before interstitials) get that order INSIDE the solver; the public arrays keep the caller's
        self.solver_order = getattr(device, "solver_order", "auto")
        self.solver_order_applied = False
        self._order_tensor = None
        # pinned staging buffers for the host<->device syncs
        self._pin = {}

    def sparsity(self, NL: int, NR: int) -> Sparsity:
        """initialize_sparsity (kmc_main.cpp:121): built once per contact size"""
        key = (NL, NR)
        if key not in self._sparsity:
            sp = Sparsity()
            check(self.ctx.lib.dkmc_initialize_sparsity(self.ctx.h, self.N_, self.nn_, _ptr(self.neigh_idx), NL, NR,
                                                        C.byref(sp)))
            self._sparsity[key] = sp
            if self.solver_order == "auto":
                from . import structures
                x, y, z = self._host_xyz
                sl = slice(NL, self.N_ - NR)
                order = structures.cell_order(x[sl], y[sl], z[sl], x0=float(x.min()))
                if not np.array_equal(order, np.arange(len(order))):
                    torch = _torch()
                    self._order_tensor = torch.from_numpy(order.astype(np.int32)).to("cuda")
                    check(self.ctx.lib.dkmc_solver_set_order(self.ctx.h, C.byref(sp), _ptr(self._order_tensor)))
                    self.solver_order_applied = True
        return self._sparsity[key]

    _SYNCED = ("site_element", "site_charge", "site_potential_boundary", "site_potential_charge", "site_temperature")

    def _pinned(self, device: Device, name: str):
        """The host array `device.<name>` re-homed (once) in page-locked memory, so that the syncs are
        plain DMA transfers with no staging copy: returns the pinned tensor that shares its storage."""
        torch = _torch()
        arr = getattr(device, name)
        ent = self._pin.get(name)
        if ent is None or ent[1] is not arr:
            t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
            view = t.numpy()
            setattr(device, name, view)
            ent = (t, view)
            self._pin[name] = ent
        return ent[0]

    def sync_HostToGPU(self, device: Device, also=()):
        """gpu_buffers.cpp:10-37.  The per-step arrays; `also` adds the per-bias-point ones
        (site_CB_edge), which the hot path never touches."""
        for name in self._SYNCED + tuple(also):
            getattr(self, name).copy_(self._pinned(device, name), non_blocking=True)
        self.T_bg.fill_(device.T_bg)

    def sync_GPUToHost(self, device: Device, also=()):
        """gpu_buffers.cpp:39-55"""
        torch = _torch()
        for name in self._SYNCED + tuple(also):
            self._pinned(device, name).copy_(getattr(self, name), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def snapshot_begin(self, device: Device) -> "Snapshot":
        """stage element / charge / potential on the device and start draining them into
        page-locked host buffers on the copy stream; returns at once.  The step may go on (and the event
        loop may mutate the site state) while the snapshot drains; Snapshot.write() produces the file
        Device::writeSnapshot would have written at this point."""
        torch = _torch()
        # the previous snapshot's writer thread formats its file from the shared staging buffers: let it finish
        last = getattr(self, "_last_snapshot", None)
        if last is not None and last._thread is not None and last._thread.is_alive():
            last._thread.join()
        if "snap" not in self._pin:
            self._pin["snap"] = (torch.empty(self.N_, dtype=torch.int32).pin_memory(),
                                 torch.empty(self.N_, dtype=torch.int32).pin_memory(),
                                 torch.empty(self.N_, dtype=torch.float64).pin_memory())
        el, q, pot = self._pin["snap"]
        check(self.ctx.lib.dkmc_snapshot_begin(self.ctx.h, self.N_, _ptr(self.site_element), _ptr(self.site_charge),
                                               _ptr(self.site_potential_boundary), _ptr(self.site_potential_charge), None,
                                               _ptr(el), _ptr(q), _ptr(pot), None))
        self._last_snapshot = Snapshot(self.ctx, device, el, q, pot)
        return self._last_snapshot

    def h2d_bytes(self) -> int:
        return self.N_ * (4 + 4 + 8 + 8 + 8)

    d2h_bytes = h2d_bytes

    def freeGPUmemory(self):
        for sp in self._sparsity.values():
            self.ctx.lib.dkmc_free_sparsity(self.ctx.h, C.byref(sp))
        self._sparsity.clear()


class Snapshot:
    """A snapshot in flight (GPUBuffers.snapshot_begin)."""

    def __init__(self, ctx: Context, device: Device, el, q, pot):
        self.ctx, self._dev, self._el, self._q, self._pot = ctx, device, el, q, pot
        self._thread = None

    def ready(self) -> bool:
        r = C.c_int(0)
        check(self.ctx.lib.dkmc_snapshot_ready(self.ctx.h, C.byref(r)))
        return bool(r.value)

    def wait(self):
        """blocks until the copies are complete; returns (element, charge, potential) host arrays
        (views of the staging buffers: valid until the next snapshot_begin)"""
        check(self.ctx.lib.dkmc_snapshot_wait(self.ctx.h))
        return self._el.numpy(), self._q.numpy(), self._pot.numpy()

    def write(self, filename: str, foldername: str = "."):
        self._write(self.wait(), filename, foldername)

    def _write(self, arrays, filename, foldername):
        import os
        el, q, pot = arrays
        d = self._dev
        write_snapshot(os.path.join(".", foldername, filename), el, d.site_x, d.site_y, d.site_z, pot)

    def write_async(self, filename: str, foldername: str = "."):
        """formats and writes the file on a host thread (the ctypes calls of the step release the GIL)"""
        import threading
        arrays = self.wait()     # the copies complete on the calling thread: only one thread talks to the context
        self._thread = threading.Thread(target=self._write, args=(arrays, filename, foldername), daemon=True)
        self._thread.start()
        return self._thread


class KMCProcess:
    """KMCProcess.h: layers, site->layer map, the KMC random stream, and the step."""

    def __init__(self, device: Device, freq: float, layers: Optional[List[Layer]] = None):
        self.random_generator = RandomNumberGenerator(RND_SEED_KMC)
        self.freq = float(freq)
        self.layers = list(layers) if layers is not None else list(DEFAULT_LAYERS)
        # KMCProcess.cpp:34-50: the LAST layer containing x wins; a site in no layer aborts
        x = device.site_x
        layer = np.full(device.N, -1, np.int32)
        for j, l in enumerate(self.layers):
            layer[(l.start_x <= x) & (x <= l.end_x)] = j
        if (layer < 0).any():
            raise ValueError(f"Site #{int(np.nonzero(layer < 0)[0][0])} is not inside the device!")
        self.site_layer = layer
        self.batch_uniforms = 4096
        self.last_events = np.zeros((0, 4), np.int32)
        self.last_info: Optional[StepInfo] = None

    def executeKMCStep(self, gpubuf: GPUBuffers, device: Device, record_events: int = 0) -> float:
        """KMCProcess.cpp:259-374 (GPU branch): returns the step time (last residence-time draw)."""
        lib = device.ctx.lib
        n_u = self.batch_uniforms
        u = self.random_generator.peek(n_u)
        info = StepInfo()
        ev = np.zeros((max(record_events, 1), 4), np.int32)
        evp = ev.ctypes.data_as(C.c_void_p) if record_events else None
        st = lib.dkmc_execute_kmc_step(
            device.ctx.h, device.N, gpubuf.nn_, _ptr(gpubuf.neigh_idx), _ptr(gpubuf.site_layer), _ptr(gpubuf.lattice),
            device.pbc, _ptr(gpubuf.T_bg), _ptr(gpubuf.freq), _ptr(gpubuf.sigma), _ptr(gpubuf.k), _ptr(gpubuf.site_x),
            _ptr(gpubuf.site_y), _ptr(gpubuf.site_z), _ptr(gpubuf.site_potential_boundary),
            _ptr(gpubuf.site_potential_charge), _ptr(gpubuf.site_element), _ptr(gpubuf.site_charge),
            u.ctypes.data_as(C.c_void_p), n_u, evp, record_events, C.byref(info))
        check(st, allow=(_capi.DKMC_ERR_RNG_EXHAUSTED,))
        self.random_generator.advance(info.n_used)
        rate_ms, loop_ms = info.rate_ms, info.loop_ms
        while st == _capi.DKMC_ERR_RNG_EXHAUSTED:
            u = self.random_generator.peek(n_u)
            st = lib.dkmc_kmc_step_continue(device.ctx.h, u.ctypes.data_as(C.c_void_p), n_u, evp, record_events,
                                            C.byref(info))
            check(st, allow=(_capi.DKMC_ERR_RNG_EXHAUSTED,))
            self.random_generator.advance(info.n_used)
            loop_ms += info.loop_ms
        info.rate_ms, info.loop_ms = rate_ms, loop_ms
        self.last_info = info
        self.last_events = ev[: min(info.n_events, record_events)].copy() if record_events else ev[:0]
        return info.event_time
