import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_step0():
    return dict(np.load(os.path.join(GOLDEN, "s_step0.npz")))


@pytest.fixture(scope="session")
def base_case(golden_step0):
    """config 0 (test_2.5nm) rebuilt from the packaged cell + the host-side RNG: positions,
    elements after makeSubstoichiometric, parameters.  Checked against the golden element array."""
    from devicekmc_b200 import structures as S
    from devicekmc_b200.host import KMCParameters, RandomNumberGenerator, O_EL, VACANCY, DEFECT, OXYGEN_DEFECT
    el, x, y, z, lat, nc = S.load_base_cell()
    p = KMCParameters()
    rng = RandomNumberGenerator(p.rnd_seed)
    atom_ind = np.nonzero((el != DEFECT) & (el != OXYGEN_DEFECT))[0]
    num_v = int(p.initial_vacancy_concentration * np.count_nonzero(el == O_EL))
    el = el.copy()
    while num_v > 0:
        loc = int(rng.getRandomNumber() * len(atom_ind))
        if el[atom_ind[loc]] == O_EL:
            el[atom_ind[loc]] = VACANCY
            num_v -= 1
    assert np.array_equal(el, golden_step0["element"].astype(np.int32)), "substoichiometric draw differs from the reference"
    return dict(element=el, x=x, y=y, z=z, lattice=lat, n_contact=nc, p=p)


def build_cpp_example(out_dir):
    """compiles examples/kmc_loop.cpp (a C++ host over the C-ABI, no reference headers) against the in-tree
    library; returns the path of the executable"""
    import subprocess
    from devicekmc_b200 import _capi
    exe = os.path.join(str(out_dir), "kmc_loop")
    libdir = os.path.dirname(_capi.LIB_PATH)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "kmc_loop.cpp"), "-o", exe,
                           "-L", libdir, "-ldkmc_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                           f"-Wl,-rpath,{libdir}"])
    return exe
