"""Momentum-space loop output ('next' row).  The reference writes parallel HDF5
(lib/loop_mugiq.cpp:530-656): dataset tree /mom_%+d_%+d_%+d/<disp tag>/<GammaName>/loop with shape [T][2].  No HDF5
library exists in this image, so a file name ending in .h5 / .hdf5 is written as a real HDF5 file by the self-contained
writer mugiq_b200/h5min.py (same tree, native double / float); any other name gets a NumPy .npz archive whose keys are
the HDF5 paths.  Tag strings are the reference's, without its group2_tag[10] truncation (disp_+z_10 no longer
collides with disp_+z_1)."""
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
    data = momentum_loop_datasets(loop)
    if str(filename).lower().endswith((".h5", ".hdf5")):
        from . import h5min
        h5min.write(filename, {"/" + k: v for k, v in data.items()})
    else:
        np.savez(filename, **data)


def write_momentum_loops_time_ranks(filename, loop, rank, world, barrier):
    """The reference's parallel write (lib/loop_mugiq.cpp:530-656) for a lattice split in t over `world` time ranks: rank 0
    lays out the HDF5 file with every [totT][2] dataset (zero-filled), `barrier()` makes it visible, then EVERY rank writes
    its own rows [rank*locT, +locT) of every dataset from its local dataMom - the hyperslab of :561-565, :624 - and a
    second barrier closes the collective.  Serial and parallel files are byte-identical."""
    from . import h5min
    local = {"/" + k: v for k, v in momentum_loop_datasets(loop).items()}   # [locT][2] each
    dt = next(iter(local.values())).dtype
    locT = next(iter(local.values())).shape[0]
    shapes = {k: (locT * world, 2) for k in local}
    if rank == 0:
        offs = h5min.skeleton(filename, shapes, dtype=dt)
    else:
        offs = h5min.offsets(shapes, dtype=dt)
    barrier()
    for k, v in local.items():
        h5min.write_rows(filename, offs[k], rank * locT, v)
    barrier()


def read_loops_file(path):
    """Reads the flat loop file the C++ host mirror writes (Loop_Mugiq::writeLoopsHDF5_Mom in
    mugiq_b200/host/src/loop_mugiq.cpp): text index + raw values.  Returns {hdf5 path: complex array [T]}."""
    with open(path, "rb") as fh:
        blob = fh.read()
    end = blob.index(b"end\n") + 4
    lines = blob[:end].decode().splitlines()
    if not lines[0].startswith("MUGIQ-B200 LOOPS v1"):
        raise ValueError(f"{path}: not a mugiq_b200 loop file")
    dt = np.float64 if lines[0].split()[-1] == "f64" else np.float32
    out = {}
    for ln in lines[1:-1]:
        _, name, T, off = ln.split()
        v = np.frombuffer(blob, dtype=dt, count=2 * int(T), offset=end + int(off)).reshape(int(T), 2)
        out[name] = v[:, 0] + 1j * v[:, 1]
    return out
