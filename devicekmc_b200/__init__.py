"""devicekmc-b200: B200-native (sm_100a) field-and-rate hot path of DeviceKMC.

`from devicekmc_b200 import Device, KMCProcess, GPUBuffers, KMCParameters` mirrors the
reference's host classes; the compute lives in lib/libdkmc_b200.so (include/dkmc.h).
"""
from .host import (DEFAULT_LAYERS, Context, Device, GPUBuffers, KMCParameters, KMCProcess, Layer,
                   RandomNumberGenerator, Snapshot, read_xyz, write_snapshot, write_xyz)
from ._capi import DkmcError, SolverOpts, SolveInfo, StepInfo, Sparsity

__all__ = ["DEFAULT_LAYERS", "Context", "Device", "GPUBuffers", "KMCParameters", "KMCProcess", "Layer",
           "RandomNumberGenerator", "Snapshot", "read_xyz", "write_snapshot", "write_xyz", "DkmcError", "SolverOpts", "SolveInfo",
           "StepInfo", "Sparsity"]
