"""Momentum-space loop output ('next' row).  The reference writes parallel HDF5
(lib/loop_mugiq.cpp:530-656); no HDF5 library exists in this image, so the same dataset tree
/mom_%+d_%+d_%+d/<disp tag>/<GammaName>/loop with shape [T][2] is stored in a NumPy .npz archive whose keys
are the HDF5 paths (a one-line h5py converter is shown in INTEGRATION.md).  Tag strings are the
reference's, without its group2_tag[10] truncation (disp_+z_10 no longer collides with disp_+z_1)."""
import numpy as np


def momentum_loop_datasets(loop):
    out = {}
    for (mom, tag, gname), arr in loop.momentum_loops().items():
        key = "mom_%+d_%+d_%+d/%s/%s/loop" % (mom[0], mom[1], mom[2], tag, gname)
        out[key] = np.stack([arr.real, arr.imag], axis=-1)
    return out


def write_momentum_loops(filename, loop):
    if not filename:
        raise ValueError("write_momentum_loops: empty filename (option --loop-mom-space-filename)")
    np.savez(filename, **momentum_loop_datasets(loop))
