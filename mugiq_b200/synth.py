"""Synthetic inputs of the shapes BASELINE.json names: random SU(3) links in host QDP
even/odd order, random normalised eigenvectors in the canonical site-major order, sigma_n = 0.01 + 0.001 n.
There is no network for real gauge configurations; QUDA's eigensolver is an external input of the path."""
import numpy as np

from .lattice import Lattice

SEED0 = 0x6D75676971  # "mugiq"


def random_gauge(L, seed=0, dtype=np.complex128):
    """[4, volume, 3, 3] complex SU(3) links, [mu][parity*volumeCB + x_cb][row][col]."""
    lat = Lattice(L)
    rng = np.random.default_rng(SEED0 ^ (seed + 1))
    n = 4 * lat.volume
    a = rng.standard_normal((n, 3, 3)) + 1j * rng.standard_normal((n, 3, 3))
    q, r = np.linalg.qr(a)
    d = np.diagonal(r, axis1=1, axis2=2)
    q = q * (d / np.abs(d))[:, None, :]          # unique QR -> Haar-distributed U(3)
    det = np.linalg.det(q)
    q = q * (det ** (-1.0 / 3.0))[:, None, None]  # det = 1
    return np.ascontiguousarray(q.reshape(4, lat.volume, 3, 3).astype(dtype))


def unit_gauge(L, dtype=np.complex128):
    lat = Lattice(L)
    g = np.zeros((4, lat.volume, 3, 3), dtype=dtype)
    g[..., 0, 0] = g[..., 1, 1] = g[..., 2, 2] = 1
    return g


def random_evecs_np(L, nEv, seed=0, dtype=np.complex128):
    """[nEv, volume, 12] complex, each normalised to 1 (orthogonality is irrelevant to the kernels)."""
    lat = Lattice(L)
    rng = np.random.default_rng(SEED0 ^ (0x9E3779B9 * (seed + 1)))
    v = rng.standard_normal((nEv, lat.volume, 12)) + 1j * rng.standard_normal((nEv, lat.volume, 12))
    v /= np.sqrt((np.abs(v) ** 2).sum(axis=(1, 2), keepdims=True))
    return np.ascontiguousarray(v.astype(dtype))


def random_evecs_torch(L, nEv, seed=0, device="cuda", dtype=None, out=None):
    """Same shape generated on the device (bench-sized sets: 200 x 16^3x32 FP64 is 5 GB); `out`: fill this tensor."""
    import torch
    dtype = dtype or torch.complex128
    lat = Lattice(L)
    gen = torch.Generator(device=device)
    gen.manual_seed((SEED0 + 7919 * (seed + 1)) & 0x7FFFFFFF)
    real = torch.float64 if dtype == torch.complex128 else torch.float32
    if out is None:
        out = torch.empty((nEv, lat.volume, 12), dtype=dtype, device=device)
    for n in range(nEv):  # one field at a time keeps the temporary small
        v = torch.randn((lat.volume, 12, 2), generator=gen, device=device, dtype=real)
        v /= v.norm()
        out[n] = torch.view_as_complex(v)
    return out


def random_gauge_slab_torch(L_global, t_slices, seed=0, device="cuda"):
    """Random SU(3) links of the given GLOBAL time-slices, generated on the device one slice at a time from a
    slice-keyed generator: every rank of a lattice-T split builds its own extended slab (interior + halo slices) of the
    SAME global field without ever holding the whole field (6.1 GB at 48^3x96).  Returns four [V4_slab, 3, 3] complex128
    tensors in even/odd order of the slab lattice (Lx, Ly, Lz, len(t_slices)); the first slice must have the parity of
    its global t (even slab offsets), as in TSplit.  Rows 1, 2 by Gram-Schmidt, row 3 = conj(row1 x row2): det = 1."""
    import torch
    Lx, Ly, Lz, T = (int(x) for x in L_global)
    V3h = Lx * Ly * Lz // 2
    nt = len(t_slices)
    out = torch.empty((4, 2, nt, V3h, 3, 3), dtype=torch.complex128, device=device)
    gen = torch.Generator(device=device)
    for i, t in enumerate(t_slices):
        gen.manual_seed((SEED0 + 104729 * (seed + 1) + 31 * (int(t) % T)) & 0x7FFFFFFF)
        a = torch.view_as_complex(torch.randn((4, 2, V3h, 2, 3, 2), generator=gen, device=device, dtype=torch.float64))
        u1 = a[..., 0, :] / a[..., 0, :].norm(dim=-1, keepdim=True)
        b = a[..., 1, :] - (u1.conj() * a[..., 1, :]).sum(-1, keepdim=True) * u1
        u2 = b / b.norm(dim=-1, keepdim=True)
        u3 = torch.linalg.cross(u1, u2).conj()
        out[:, :, i] = torch.stack([u1, u2, u3], dim=-2)
    return [out[mu].reshape(2 * nt * V3h, 3, 3) for mu in range(4)]


def sigmas(nEv):
    return 0.01 + 0.001 * np.arange(nEv, dtype=np.float64)


ONE_HOP_ENTRIES = "+x:1;-x:1;+y:1;-y:1;+z:1;-z:1;+t:1;-t:1"  # config 2 of BASELINE.json (nLoop = 9)
UP_TO_4_ENTRIES = "+x:1,4;-x:1,4;+y:1,4;-y:1,4;+z:1,4;-z:1,4;+t:1,4;-t:1,4"  # config 3 (nLoop = 33)
