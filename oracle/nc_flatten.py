"""oracle/nc_flatten.py -- TEST INFRASTRUCTURE, not product code.

Writes <file>.ncflat next to (or at `out`) a NetCDF-4 / HDF5 file: every dataset in its stored type, in the flat container
that oracle/shim/netcdf.h serves to the reference's NetCDF_Reader (see that header for the format).  libnetcdf / libhdf5 are
not in this image; the bytes are read with the repo's own dependency-free HDF5 subset reader (axisem3d_b200/h5lite.py).

    python oracle/nc_flatten.py tests/golden/AxiSEM_prem_ani_one_crust_50.e /tmp/run/input/AxiSEM_prem_ani_one_crust_50.e.ncflat
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from axisem3d_b200 import h5lite  # noqa: E402

_TYPES = {"i1": "i1", "u1": "i1", "S1": "c1", "i4": "i4", "u4": "i4", "i8": "i8", "u8": "i8", "f4": "f4", "f8": "f8"}


def flatten(path, out=None):
    f = h5lite.File(path)
    out = out or path + ".ncflat"
    items = []
    for name in sorted(f.keys()):
        try:
            a = np.ascontiguousarray(f[name].read())
        except h5lite.H5Error:
            continue                       # dimension scales without data etc.
        key = a.dtype.str.lstrip("<|=")
        if key.startswith("S") and key != "S1":
            # fixed-length strings: [n][len] characters, NUL padded (what nc_get_var_text hands back)
            a = np.frombuffer(a.tobytes(), dtype="S1").reshape(a.shape + (a.dtype.itemsize,))
            key = "S1"
        if key not in _TYPES:
            continue
        items.append((name, _TYPES[key], a))
    with open(out, "wb") as g:
        g.write(b"NCFLAT1\n%d\n" % len(items))
        for name, t, a in items:
            g.write(("%s %s %d %s %d\n" % (name, t, a.ndim, " ".join(str(int(n)) for n in a.shape), a.nbytes)).encode())
            g.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())
    return out


def read_flat(path):
    """The container back as {name: ndarray} (used by the tests to read what the reference wrote through the stand-in)."""
    types = {"i1": np.int8, "c1": "S1", "i4": np.int32, "i8": np.int64, "f4": np.float32, "f8": np.float64}
    res = {}
    with open(path, "rb") as g:
        assert g.readline().startswith(b"NCFLAT1")
        for _ in range(int(g.readline())):
            parts = g.readline().split()
            name, t, nd = parts[0].decode(), parts[1].decode(), int(parts[2])
            shape = tuple(int(x) for x in parts[3:3 + nd])
            raw = g.read(int(parts[3 + nd]))
            res[name] = np.frombuffer(raw, dtype=types[t]).reshape(shape) if nd else np.frombuffer(raw, dtype=types[t])
    return res


if __name__ == "__main__":
    print(flatten(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None))
