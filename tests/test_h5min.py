"""mugiq_b200/h5min.py: the self-contained HDF5 writer of the momentum-space loop file and its reader.
The reader is checked against a REAL HDF5 file written by the HDF5 library (a MATLAB 7.3 file in scipy's test data, the
only HDF5 file in this image), the writer against the reader and against the structures of that file."""
import glob
import os
import struct

import numpy as np
import pytest

from mugiq_b200 import h5min


def _real_hdf5_file():
    try:
        import scipy.io
    except Exception:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


def test_reader_on_a_file_written_by_the_hdf5_library():
    path = _real_hdf5_file()
    if path is None:
        pytest.skip("scipy's MATLAB 7.3 test file is not available")
    out = h5min.read(path)
    assert list(out) == ["/testdouble"]
    # MATLAB: testdouble = 0:pi/4:2*pi
    assert out["/testdouble"].dtype == np.float64 and np.allclose(out["/testdouble"].ravel(), np.arange(9) * np.pi / 4, atol=1e-15)


def test_writer_uses_the_encodings_of_the_library():
    """Field-by-field comparison with the real file for everything that does not depend on addresses."""
    path = _real_hdf5_file()
    if path is None:
        pytest.skip("scipy's MATLAB 7.3 test file is not available")
    real = h5min._Reader(open(path, "rb").read())
    mine = h5min._Reader(h5min.dumps({"/testdouble": (np.arange(9) * np.pi / 4).reshape(9, 1)}))
    def dataset_msgs(r):
        m = dict(r.messages(r.root_oh))
        tree, heap = struct.unpack("<QQ", m[0x0011][:16])
        (name, oh), = r.links(tree, heap)
        assert name == "testdouble"
        return dict(r.messages(oh)), tree, heap
    a, ta, ha = dataset_msgs(real)
    b, tb, hb = dataset_msgs(mine)
    assert a[0x0003][:20] == b[0x0003][:20]           # IEEE double, little endian: identical datatype message
    assert a[0x0001] == b[0x0001]                     # dataspace [9][1], version 1
    assert real.at(real.root_oh, 2) == mine.at(mine.root_oh, 2)                        # object header version
    assert real.at(ta, 8) == mine.at(tb, 8) and real.at(ta + 24, 8) == mine.at(tb + 24, 8)  # B-tree node header, first key
    assert real.at(ha, 8) == mine.at(hb, 8)                                             # local heap signature + version
    assert real.internal_k == mine.internal_k == 16 and real.leaf_k == mine.leaf_k == 4
    # message flags (bit 0 = constant) of the messages both files have
    def flags(r, oh):
        out, pos = {}, oh + 16
        for _ in range(struct.unpack("<H", r.at(oh + 2, 2))[0]):
            t, sz, fl = struct.unpack("<HHB", r.at(pos, 5))
            out[t] = fl
            pos += 8 + sz
        return out
    (_, oh_a), = real.links(ta, ha)
    (_, oh_b), = mine.links(tb, hb)
    fa, fb = flags(real, oh_a), flags(mine, oh_b)
    assert all(fa[t] == fb[t] for t in (0x0001, 0x0003, 0x0005, 0x0008))
    assert flags(real, real.root_oh)[0x0011] == flags(mine, mine.root_oh)[0x0011]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_loop_file_round_trip(tmp_path, dtype):
    from mugiq_b200.params import GAMMA_NAMES
    rng = np.random.default_rng(3)
    moms = [(0, 0, 0), (-1, 0, 1), (1, -1, 0), (0, 2, -2)]
    tags = ["disp_0", "disp_+x_1", "disp_-t_2", "disp_+z_10", "disp_+z_1"]   # no group2_tag[10] truncation
    T = 12
    want = {}
    for m in moms:
        for tag in tags:
            for g in range(16):
                want["/mom_%+d_%+d_%+d/%s/%s/loop" % (m[0], m[1], m[2], tag, GAMMA_NAMES[g])] = rng.standard_normal((T, 2)).astype(dtype)
    f = tmp_path / "loops.h5"
    n = h5min.write(str(f), want)
    blob = f.read_bytes()
    assert len(blob) == n and blob[:8] == h5min.SIG and struct.unpack_from("<Q", blob, 40)[0] == n  # end-of-file address
    got = h5min.loads(blob)
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].dtype == dtype and np.array_equal(got[k], want[k]), k
    assert h5min.dumps(want) == blob  # deterministic output


def test_many_links_in_one_group_and_errors():
    many = {f"/g/d{i:04d}": np.full((2, 2), float(i)) for i in range(300)}
    got = h5min.loads(h5min.dumps(many))
    assert len(got) == 300 and all(got[k][0, 0] == float(k[-4:]) for k in got)
    with pytest.raises(TypeError):
        h5min.dumps({"/a": np.zeros(3, dtype=np.int32)})
    with pytest.raises(ValueError):
        h5min.dumps({"/a": np.zeros(3), "/a/b": np.zeros(3)})


def test_cpp_writer_is_byte_identical(tmp_path):
    """mugiq_b200/host/src/h5min.hpp (what the C++ Loop_Mugiq::writeLoopsHDF5_Mom uses) against the Python writer."""
    import shutil
    import subprocess
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    hpp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mugiq_b200", "host", "src", "h5min.hpp")
    src = tmp_path / "t.cpp"
    src.write_text('#include "%s"\n' % hpp + r'''
int main(int, char **argv) {
  h5min::File f;
  double a[24]; float b[8];
  for (int i = 0; i < 24; i++) a[i] = 0.25 * i - 1.0;
  for (int i = 0; i < 8; i++) b[i] = 1.5f * i;
  const char *g[3] = {"g5g4", "1", "g1g2"};
  for (int m = -1; m <= 1; m++)
    for (int k = 0; k < 3; k++) {
      char p[128];
      snprintf(p, sizeof(p), "/mom_%+d_%+d_%+d/disp_+z_%d/%s/loop", m, 0, -m, k == 2 ? 10 : 1, g[k]);
      f.addDataset(p, {12, 2}, 8, a, sizeof(a));
    }
  f.addDataset("/single/loop", {4, 2}, 4, b, sizeof(b));
  return f.write(argv[1]) ? 0 : 1;
}
''')
    exe = tmp_path / "t"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-o", str(exe), str(src)])
    out = tmp_path / "cpp.h5"
    subprocess.check_call([str(exe), str(out)])
    blob = out.read_bytes()
    h = h5min.loads(blob)
    assert len(h) == 10 and h["/single/loop"].dtype == np.float32 and h["/mom_+1_+0_-1/disp_+z_10/g1g2/loop"].shape == (12, 2)
    assert np.array_equal(h["/mom_-1_+0_+1/disp_+z_1/1/loop"].ravel(), 0.25 * np.arange(24) - 1.0)
    assert h5min.dumps(h) == blob


def test_time_ranks_write_their_rows_into_one_file(tmp_path):
    """The reference's parallel write (MPI-IO + hyperslabs, lib/loop_mugiq.cpp:561-572,624): rank 0 lays the file out, every
    time rank writes its rows [rank*locT, +locT) of every [totT][2] dataset.  The result must be byte-identical to the
    serial file, whatever the order the ranks write in."""
    from mugiq_b200 import h5min
    rng = np.random.default_rng(3)
    T, world = 12, 3
    names = ["/mom_%+d_+0_+0/disp_%s/%s/loop" % (p, d, g) for p in (-1, 0, 1) for d in ("0", "+z_1", "-t_10") for g in ("G0", "G5", "G15")]
    full = {n: rng.standard_normal((T, 2)) for n in names}
    serial = tmp_path / "serial.h5"
    h5min.write(serial, full)
    par = tmp_path / "parallel.h5"
    shapes = {n: (T, 2) for n in names}
    offs = h5min.skeleton(par, shapes)
    assert offs == h5min.offsets(shapes)                      # every rank can derive the layout itself
    assert all(np.array_equal(v, np.zeros((T, 2))) for v in h5min.read(par).values())
    locT = T // world
    for rank in (2, 0, 1):
        for n in names:
            h5min.write_rows(par, offs[n], rank * locT, full[n][rank * locT:(rank + 1) * locT])
    assert par.read_bytes() == serial.read_bytes()
    got = h5min.read(par)
    assert set(got) == set(names) and all(np.array_equal(got[n], full[n]) for n in names)
    # single precision
    f32 = {n: v.astype(np.float32) for n, v in full.items()}
    h5min.write(serial, f32)
    offs = h5min.skeleton(par, shapes, dtype=np.float32)
    for rank in range(world):
        for n in names:
            h5min.write_rows(par, offs[n], rank * locT, f32[n][rank * locT:(rank + 1) * locT])
    assert par.read_bytes() == serial.read_bytes()
