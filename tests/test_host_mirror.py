"""The C++ host-side mirror of the reference interface (mugiq_b200/host: Loop_Mugiq, Displace, computeLoop, the wrapper
functions and the loop_driver executable standing in for the reference's tests/loop.cpp).  CPU part: it builds, links
against the C-ABI library, parses the reference's option grammar and aborts like errorQuda on bad input.  GPU part:
the driver's results equal the oracle's for site-major, FLOAT2 and FLOAT4 eigenvector orders."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err, TOL_F64, TOL_F32

HOST = os.path.join(ROOT, "mugiq_b200", "host")
DRIVER = os.path.join(HOST, "build", "loop_driver")


@pytest.fixture(scope="module")
def driver():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mugiq_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", HOST])
    assert os.path.exists(DRIVER)
    return DRIVER


def run(driver, *args):
    return subprocess.run([driver, *[str(a) for a in args]], capture_output=True, text=True)


def test_host_library_exports_reference_interface(driver):
    out = subprocess.run(["nm", "-DC", os.path.join(ROOT, "mugiq_b200", "lib", "libmugiq_host.so")], capture_output=True,
                         text=True).stdout
    for sym in ["Loop_Mugiq<double, (QudaFieldOrder_s)2>::computeCoarseLoop()",
                "Loop_Mugiq<float, (QudaFieldOrder_s)4>::writeLoopsHDF5()",
                "Displace<double, (QudaFieldOrder_s)2>::Displace(MugiqLoopParam_s*, quda::ColorSpinorField*, QudaPrecision_s)",
                "void computeLoop<double>(QudaMultigridParam_s, QudaEigParam_s, MugiqLoopParam_s, MuGiqBool_s, MuGiqBool_s)",
                "void performLoopContraction<double, (QudaFieldOrder_s)2>(std::complex<double>*, quda::ColorSpinorField*, "
                "quda::ColorSpinorField*, double)",
                "void convertIdxOrder_mapGamma<float>(std::complex<float>*, std::complex<float> const*, int, int, int, int, "
                "int const*)",
                "void createPhaseMatrixGPU<double>(std::complex<double>*, int const*, long long, int, int, int const*, int const*)",
                "void performCovariantDisplacementVector<double, (QudaFieldOrder_s)4>"]:
        assert sym in out, sym


def test_driver_parses_reference_option_grammar(driver, tmp_path):
    mom = tmp_path / "mom.txt"
    mom.write_text("0 0 0\n1 0 0\n\n0 -1 1\n")
    r = run(driver, "--parse-only", "--loop-do-nonlocal", "yes", "--displace-entry-string", "+z:1,8;-x:3;+t:5,2",
            "--loop-do-momproj", "yes", "--momenta-filename", mom, "--loop-ft-sign", "plus")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "nonlocal 1 entries 3"
    assert lines[1:4] == ["entry 0 +z 1 8", "entry 1 -x 3 3", "entry 2 +t 5 2"]
    assert "momproj 1 Nmom 3 ftsign 1" in lines
    assert lines[-1] == "mom 0 -1 1"


@pytest.mark.parametrize("entry,msg", [("+z:1:8", "Wrong format"), ("+z:1,2,3", "Wrong format"), ("+z:a", "Wrong format"),
                                       ("", "--displace-entry-string is not set")])
def test_driver_aborts_like_errorQuda(driver, entry, msg):
    r = run(driver, "--parse-only", "--loop-do-nonlocal", "yes", "--displace-entry-string", entry)
    assert r.returncode != 0 and msg in r.stderr


def test_driver_rejects_bad_momenta_file(driver, tmp_path):
    mom = tmp_path / "mom.txt"
    mom.write_text("0 0 0\n1 x 0\n")
    r = run(driver, "--parse-only", "--loop-do-momproj", "yes", "--momenta-filename", mom)
    assert r.returncode != 0 and "Incorrect file format in Line 1" in r.stderr


def _inputs(tmp_path, L, nEv, prec):
    from mugiq_b200 import synth
    cdt = np.complex128 if prec == "double" else np.complex64
    ev = synth.random_evecs_np(L, nEv, seed=51).astype(cdt)
    U = synth.random_gauge(L, seed=51).astype(cdt)
    sig = synth.sigmas(nEv)
    ev.tofile(tmp_path / "ev.bin")
    U.tofile(tmp_path / "u.bin")
    sig.tofile(tmp_path / "sig.bin")
    return ev, U, sig


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["site", "float2", "float4"])
@pytest.mark.parametrize("prec", ["double", "single"])
def test_driver_matches_oracle(driver, oracle, tmp_path, order, prec):
    from oracle import numpy_check as npc
    from mugiq_b200.h5lite import read_loops_file
    from mugiq_b200.params import momenta_up_to, GAMMA_NAMES
    L, nEv = (4, 4, 4, 8), 6
    ev, U, sig = _inputs(tmp_path, L, nEv, prec)
    mom = momenta_up_to(1)
    (tmp_path / "mom.txt").write_text("".join(f"{p[0]} {p[1]} {p[2]}\n" for p in mom))
    entries = [(2, 1, 1, 3), (0, 0, 2, 2), (3, 0, 1, 1), (3, 1, 1, 1)]
    r = run(driver, "--dim", *L, "--prec", prec, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file",
            tmp_path / "sig.bin", "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string",
            "+z:1,3;-x:2;-t:1;+t:1", "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--loop-ft-sign",
            "minus", "--loop-write-mom-space", "yes", "--loop-mom-space-filename", tmp_path / "loops.dat", "--field-order", order,
            "--dump-pos", tmp_path / "pos.bin", "--dump-mom", tmp_path / "mom.bin")
    assert r.returncode == 0, r.stderr
    cdt = np.complex128 if prec == "double" else np.complex64
    tol = TOL_F64 if prec == "double" else TOL_F32
    ref = oracle.compute_loop(ev.astype(np.complex128), sig, U.astype(np.complex128), entries, L)
    pos = np.fromfile(tmp_path / "pos.bin", dtype=cdt).reshape(ref.shape)
    assert rel_err(pos, ref) < tol
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    dm = np.fromfile(tmp_path / "mom.bin", dtype=cdt).reshape(ref_mom.shape)
    assert rel_err(dm, ref_mom) < tol
    loops = read_loops_file(tmp_path / "loops.dat")
    assert len(loops) == len(mom) * ref.shape[0] * 16
    im = mom.index([0, 1, 0])
    got = loops["/mom_+0_+1_+0/disp_+z_3/" + GAMMA_NAMES[7] + "/loop"]
    assert np.abs(got - ref_mom[im, 7 + 16 * 3, :]).max() < tol * np.abs(ref_mom).max()
    if order == "site":
        # the same run with an .h5 file name: a real HDF5 file (h5min.hpp), byte-identical to what the Python writer
        # (mugiq_b200/h5min.py) produces for the same datasets, and readable by the independent reader
        from mugiq_b200 import h5min
        r = run(driver, "--dim", *L, "--prec", prec, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file",
                tmp_path / "sig.bin", "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string",
                "+z:1,3;-x:2;-t:1;+t:1", "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--loop-ft-sign",
                "minus", "--loop-write-mom-space", "yes", "--loop-mom-space-filename", tmp_path / "loops.h5")
        assert r.returncode == 0, r.stderr
        blob = (tmp_path / "loops.h5").read_bytes()
        h = h5min.loads(blob)
        rdt = np.float64 if prec == "double" else np.float32
        assert set(h) == set(loops) and all(v.shape == (L[3], 2) and v.dtype == rdt for v in h.values())
        assert all(np.array_equal(h[k][:, 0] + 1j * h[k][:, 1], loops[k]) for k in loops)
        assert h5min.dumps(h) == blob


@pytest.mark.gpu
@pytest.mark.parametrize("peer", ["no", "yes"])
def test_driver_eigenvector_shards_over_nccl(driver, oracle, tmp_path, peer):
    """Two driver processes, one per GPU, each with its eigenvector shard; the loop buffer is summed chunk by chunk under
    the kernels, through NCCL or (--peer-reduce yes) by copy-engine pushes over peer-mapped buffers (comm_mugiq.h).
    Every rank ends with the full result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not put two ranks on one device)")
    from oracle import numpy_check as npc
    from mugiq_b200.params import momenta_up_to
    L, nEv = (4, 4, 4, 8), 7
    ev, U, sig = _inputs(tmp_path, L, nEv, "double")
    mom = momenta_up_to(1)
    (tmp_path / "mom.txt").write_text("".join(f"{p[0]} {p[1]} {p[2]}\n" for p in mom))
    entries = [(2, 1, 1, 2), (2, 0, 1, 2), (0, 0, 1, 1)]
    procs = []
    for rank in range(2):
        args = [driver, "--dim", *L, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file", tmp_path / "sig.bin",
                "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string", "+z:1,2;-z:1,2;-x:1",
                "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--comm-size", 2, "--comm-rank", rank,
                "--comm-id-file", tmp_path / "nccl.id", "--device", rank, "--dump-pos", tmp_path / f"pos{rank}.bin", "--dump-mom",
                tmp_path / f"mom{rank}.bin", "--peer-reduce", peer, "--verbosity", "verbose"]
        procs.append(subprocess.Popen([str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err
        assert ("position-space sum over peer-mapped buffers" in out) == (peer == "yes")
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    for rank in range(2):
        pos = np.fromfile(tmp_path / f"pos{rank}.bin", dtype=np.complex128).reshape(ref.shape)
        dm = np.fromfile(tmp_path / f"mom{rank}.bin", dtype=np.complex128).reshape(ref_mom.shape)
        assert rel_err(pos, ref) < TOL_F64 and rel_err(dm, ref_mom) < TOL_F64


def _stitch_time_slabs(parts, ref_shape, L, world):
    """[rank] local dataPos [nLoop, 16, V4_loc] (even/odd order of the slab) -> global [nLoop, 16, V4]"""
    V3h, Tl = L[0] * L[1] * L[2] // 2, L[3] // world
    return np.concatenate([p.reshape(ref_shape[0], 16, 2, Tl, V3h) for p in parts], axis=3).reshape(ref_shape)


@pytest.mark.gpu
@pytest.mark.parametrize("entry_str,entries", [("+t:1,2;-t:1,2;+x:1;-y:2", [(3, 1, 1, 2), (3, 0, 1, 2), (0, 1, 1, 1), (1, 0, 2, 2)]),
                                               ("-t:1;+z:1", [(3, 0, 1, 1), (2, 1, 1, 1)])])
def test_driver_time_split_single_rank(driver, oracle, tmp_path, entry_str, entries):
    """C++ Loop_Mugiq on a lattice 'partitioned' in t over one rank: extended slabs in a peer allocation, halo pushes into
    itself (periodic), interior-only kernels, loop-buffer halo of the derived minus-t loops - against the oracle."""
    from oracle import numpy_check as npc
    from mugiq_b200.params import momenta_up_to
    L, nEv = (4, 4, 2, 8), 5
    ev, U, sig = _inputs(tmp_path, L, nEv, "double")
    mom = momenta_up_to(1)
    (tmp_path / "mom.txt").write_text("".join(f"{p[0]} {p[1]} {p[2]}\n" for p in mom))
    r = run(driver, "--dim", *L, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file", tmp_path / "sig.bin",
            "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string", entry_str,
            "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--tsplit", 1,
            "--dump-pos", tmp_path / "pos.bin", "--dump-mom", tmp_path / "mom.bin")
    assert r.returncode == 0, r.stderr
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    assert rel_err(np.fromfile(tmp_path / "pos.bin", dtype=np.complex128).reshape(ref.shape), ref) < TOL_F64
    assert rel_err(np.fromfile(tmp_path / "mom.bin", dtype=np.complex128).reshape(ref_mom.shape), ref_mom) < TOL_F64


@pytest.mark.gpu
def test_driver_time_split_over_two_gpus(driver, oracle, tmp_path):
    """Two driver processes, one per GPU, each with its time slab: CUDA-IPC mapped slabs, halo slices pushed over NVLink,
    NCCL for the flags and the time gather.  The stitched position-space buffer and the gathered momentum-space buffer
    (on both ranks) equal the oracle's on the global lattice."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import numpy_check as npc
    from mugiq_b200.params import momenta_up_to
    L, nEv = (4, 4, 2, 8), 5
    ev, U, sig = _inputs(tmp_path, L, nEv, "double")
    mom = momenta_up_to(1)
    (tmp_path / "mom.txt").write_text("".join(f"{p[0]} {p[1]} {p[2]}\n" for p in mom))
    entries = [(3, 1, 1, 2), (3, 0, 1, 2), (0, 1, 1, 1), (3, 0, 1, 1)]
    procs = []
    for rank in range(2):
        args = [driver, "--dim", *L, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file", tmp_path / "sig.bin",
                "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string", "+t:1,2;-t:1,2;+x:1;-t:1",
                "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--tsplit", 2, "--comm-size", 2,
                "--comm-rank", rank, "--comm-id-file", tmp_path / "nccl.id", "--device", rank, "--dump-pos",
                tmp_path / f"pos{rank}.bin", "--dump-mom", tmp_path / f"mom{rank}.bin"]
        procs.append(subprocess.Popen([str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    parts = [np.fromfile(tmp_path / f"pos{r}.bin", dtype=np.complex128) for r in range(2)]
    assert rel_err(_stitch_time_slabs(parts, ref.shape, L, 2), ref) < TOL_F64
    for rank in range(2):
        dm = np.fromfile(tmp_path / f"mom{rank}.bin", dtype=np.complex128).reshape(ref_mom.shape)
        assert rel_err(dm, ref_mom) < TOL_F64


@pytest.mark.gpu
def test_driver_public_entry_point_and_fatal_errors(driver, tmp_path):
    """computeLoop<Float>() through the registry (no dumps), then the reference's fatal paths: position-space writing is
    'Not supported yet', a precision mismatch aborts."""
    L, nEv = (4, 4, 4, 4), 3
    _inputs(tmp_path, L, nEv, "double")
    base = ["--dim", *L, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file", tmp_path / "sig.bin", "--gauge-file",
            tmp_path / "u.bin"]
    (tmp_path / "mom.txt").write_text("0 0 0\n")
    r = run(driver, *base, "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--loop-write-mom-space", "yes",
            "--loop-mom-space-filename", tmp_path / "l.dat")
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(tmp_path / "l.dat") > 16 * 4 * 16
    r = run(driver, *base, "--loop-write-pos-space", "yes")
    assert r.returncode != 0 and "Not supported yet" in r.stderr
    r = run(driver, *base, "--prec", "single")
    assert r.returncode != 0  # evecs file has double-precision size


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["site", "float2"])
def test_driver_streamed_eigenvectors_match_oracle(driver, oracle, tmp_path, order):
    """The producer hook of the C++ mirror (Eigsolve_Mugiq::setEvecProducer, standing in for prolongateEvec,
    /root/reference/lib/loop_mugiq.cpp:276-319,482): eigenvectors stay in host memory and travel through the library's feed
    (staging batches of 4 out of 7 eigenvectors) instead of living on the device."""
    from oracle import numpy_check as npc
    from mugiq_b200.params import momenta_up_to
    L, nEv = (4, 4, 4, 8), 7
    ev, U, sig = _inputs(tmp_path, L, nEv, "double")
    mom = momenta_up_to(1)
    (tmp_path / "mom.txt").write_text("".join(f"{p[0]} {p[1]} {p[2]}\n" for p in mom))
    entries = [(2, 1, 1, 3), (0, 0, 2, 2), (3, 0, 1, 1), (3, 1, 1, 1)]
    r = run(driver, "--dim", *L, "--n-ev", nEv, "--evecs-file", tmp_path / "ev.bin", "--sigma-file", tmp_path / "sig.bin",
            "--gauge-file", tmp_path / "u.bin", "--loop-do-nonlocal", "yes", "--displace-entry-string", "+z:1,3;-x:2;-t:1;+t:1",
            "--loop-do-momproj", "yes", "--momenta-filename", tmp_path / "mom.txt", "--field-order", order, "--stream-evecs", "yes",
            "--stream-batch", 4, "--dump-pos", tmp_path / "pos.bin", "--dump-mom", tmp_path / "mom.bin")
    assert r.returncode == 0, r.stderr
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    assert rel_err(np.fromfile(tmp_path / "pos.bin", dtype=np.complex128).reshape(ref.shape), ref) < TOL_F64
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    assert rel_err(np.fromfile(tmp_path / "mom.bin", dtype=np.complex128).reshape(ref_mom.shape), ref_mom) < TOL_F64


@pytest.mark.gpu
def test_driver_bench_mode_prints_the_e2e_fields(driver):
    """loop_driver --bench: end-to-end timing of the C++ front end on synthetic host data (what bench.py reports as
    e2e_cpp); the run checks its own result (sum_x T_1 = sum 1/sigma) and exits non-zero if it is off."""
    import json
    r = run(driver, "--bench", "--dim", 8, 4, 4, 8, "--n-ev", 7, "--displace-entry-string", "+x:1;-x:1;-t:1,2", "--bench-p2max", 1,
            "--bench-steps", 2, "--bench-batch", 3)
    assert r.returncode == 0, r.stderr + r.stdout
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["steps"] == 2 and d["nLoop"] == 5 and d["Nmom"] == 7 and d["value"] > 0 and d["checksum_rel_err"] < 1e-10
    assert d["h2d_bytes_per_step"] == 7 * 8 * 4 * 4 * 8 * 192 + 4 * 8 * 4 * 4 * 8 * 144
    assert d["d2h_bytes_per_step"] == (5 * 16 * 8 * 4 * 4 * 8 + 16 * 7 * 8 * 5) * 16
