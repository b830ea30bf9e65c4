"""User-side helpers that mirror ``pytdscf.util`` where the hot path's outputs are concerned.

``read_nc`` has the call signature and return value of the reference's reader (pytdscf/util/read_nc.py:4-24: a dict with
``"time"`` and one complex array of shape (step, d, d, ...) per requested key) and reads the ``reduced_density.nc`` this package
writes (``simulator_cls._write_reduced_density_nc``): NetCDF-3 with a trailing ``complex`` dimension of length 2 in place of the
reference's NETCDF4 compound (real, imag) type, because neither the netCDF4 package nor HDF5 exists in the image."""
from __future__ import annotations

import numpy as np


def read_nc(filename: str, sites: list[tuple[int, ...]]) -> dict:
    from scipy.io import netcdf_file

    data: dict = {}
    with netcdf_file(filename, "r", mmap=False) as f:
        data["time"] = np.array(f.variables["time"][:])
        for key in sites:
            varname = f"rho_{tuple(key)}_0"
            if varname not in f.variables:
                raise ValueError(f"Density data for site {key} {varname=} not found in {filename}")
            raw = np.array(f.variables[varname][:])
            data[key] = raw[..., 0] + 1.0j * raw[..., 1]
    return data
