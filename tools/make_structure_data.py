"""Converts the reference's shipped input structures (DATA, not code) into the compact .npz files under
devicekmc_b200/data/ that the benchmark configurations 2 and 3 use (BASELINE.json configs).
Run in the build container, where /root/reference exists; the .npz files are committed and travel to the GPU box.
    python tools/make_structure_data.py
Sites keep the order of the file — the reference's order (contacts first / last, lattice atoms before
interstitials, reorder_boundary.py:113-124); n_contact = num_atoms_contact of the structure's parameters.txt, the
contact size of the reference's CPU branch (potential_solver.cpp:271,294) — the semantics the oracle follows.  Its
GPU branch passes num_atoms_first_layer instead (potential_solver.cpp:240-241; kept as n_first_layer): on the
crossbar that leaves two of the four electrode lines floating (5 761 and 7 920 metal sites tied by high_G, held
only through low_G); on the 2 x 2 tile with that contact size the GPU solve used up 20 000 iterations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devicekmc_b200.host import read_xyz  # noqa: E402

REF = "/root/reference/structures"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "devicekmc_b200", "data")
ITEMS = {
    # name: (xyz file, lattice, num_atoms_contact, num_atoms_first_layer, V_switch)  — from the parameters.txt next to the file
    "device_7.5nm": ("single_devices/7.5nm_potential/reordered_device_7.5.xyz", (108.984050, 76.725, 76.725), 12960, 1296, 10.0),
    "crossbar_10nm_5pitch": ("crossbars/10nm_5pitch/reordered_crossbar_10_5.xyz", (108.98, 102.3, 102.3), 11520, 144, 15.0),
}
for name, (rel, lattice, nc, nfl, vd) in ITEMS.items():
    el, x, y, z = read_xyz(os.path.join(REF, rel))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), element=el.astype(np.int8), xyz=np.stack([x, y, z], 1),
                        lattice=np.array(lattice), n_contact=np.int32(nc), n_first_layer=np.int32(nfl), V_switch=np.float64(vd))
    print(name, len(x), "sites", os.path.getsize(os.path.join(OUT, name + ".npz")) >> 10, "KiB")
