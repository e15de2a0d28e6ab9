"""Host-side lattice geometry: QUDA even/odd site order <-> coordinates <-> lexicographic order.

Restates the index conventions the reference inherits from QUDA (getCoords / linkIndex, used at
/root/reference/lib/mugiq_displace_kernels.cu:168 and lib/mugiq_util_kernels.cu:75) for host bookkeeping,
synthetic-input generation and result slicing.  No arithmetic of the hot path lives here.
"""
import numpy as np


class Lattice:
    def __init__(self, L):
        L = tuple(int(x) for x in L)
        if len(L) != 4 or any(x < 1 for x in L):
            raise ValueError(f"lattice extents must be 4 positive ints, got {L}")
        if L[0] % 2:
            raise ValueError("L[0] must be even (even/odd site order)")
        self.L = L
        self.V3 = L[0] * L[1] * L[2]
        self.volume = self.V3 * L[3]
        self.volumeCB = self.volume // 2

    # -- coordinates of every site in even/odd order: array [volume, 4] (x,y,z,t) ------------------------
    def coords_eo(self):
        Lx, Ly, Lz, _ = self.L
        cb = np.arange(self.volumeCB, dtype=np.int64)
        out = np.empty((2, self.volumeCB, 4), dtype=np.int64)
        for parity in (0, 1):
            za = cb // (Lx // 2)
            zb = za // Ly
            y = za - zb * Ly
            t = zb // Lz
            z = zb - t * Lz
            x = 2 * cb + ((y + z + t + parity) & 1) - za * Lx
            out[parity, :, 0], out[parity, :, 1], out[parity, :, 2], out[parity, :, 3] = x, y, z, t
        return out.reshape(self.volume, 4)

    def lex_of_eo(self):
        """lexicographic index x + Lx*(y + Ly*(z + Lz*t)) of every even/odd-ordered site."""
        c = self.coords_eo()
        Lx, Ly, Lz, _ = self.L
        return c[:, 0] + Lx * (c[:, 1] + Ly * (c[:, 2] + Lz * c[:, 3]))

    def eo_of_lex(self):
        """inverse permutation: even/odd index of every lexicographic site."""
        lex = self.lex_of_eo()
        inv = np.empty_like(lex)
        inv[lex] = np.arange(self.volume)
        return inv

    def eo_index(self, x, y, z, t):
        Lx, Ly, Lz, _ = self.L
        lex = x + Lx * (y + Ly * (z + Lz * t))
        return (lex >> 1) + ((x + y + z + t) & 1) * self.volumeCB

    def neighbour_eo(self, direction, sign):
        """even/odd index of x + mu (sign=1) or x - mu (sign=0) for every even/odd-ordered site (periodic)."""
        c = self.coords_eo().copy()
        c[:, direction] = (c[:, direction] + (1 if sign else -1)) % self.L[direction]
        return self.eo_index(c[:, 0], c[:, 1], c[:, 2], c[:, 3])
