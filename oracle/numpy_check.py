"""Independent numpy restatement of the loop path — TEST INFRASTRUCTURE ONLY.

A second, structurally different derivation used to pin oracle/mugiq_oracle.cpp (the reference has no golden
vectors): dense 4x4 gamma matrices built from the textbook DeGrand-Rossi g1..g4 and
G(n) = g1^n0 g2^n1 g3^n2 g4^n3 (/root/reference/include/gamma.h:24-27), fields held as lexicographic
[t,z,y,x,...] arrays and displaced with np.roll, Fourier phases from np.exp.  Nothing here shares code or
tables with the C++ oracle or the CUDA kernels."""
import numpy as np

I = 1j
# DeGrand-Rossi basis (the basis QUDA/QDP call "DeGrand-Rossi"; include/gamma.h:23)
G1 = np.array([[0, 0, 0, I], [0, 0, I, 0], [0, -I, 0, 0], [-I, 0, 0, 0]])
G2 = np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=complex)
G3 = np.array([[0, 0, I, 0], [0, 0, 0, -I], [-I, 0, 0, 0], [0, I, 0, 0]])
G4 = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=complex)


def gamma_dense():
    """[16,4,4] with G(n) = g1^n0 g2^n1 g3^n2 g4^n3, n = n0 + 2 n1 + 4 n2 + 8 n3."""
    out = np.zeros((16, 4, 4), dtype=complex)
    for n in range(16):
        m = np.eye(4, dtype=complex)
        for bit, g in enumerate((G1, G2, G3, G4)):
            if (n >> bit) & 1:
                m = m @ g
        out[n] = m
    return out


def lex_coords(L):
    """eo-ordered site -> (x,y,z,t), computed from first principles: parity = (x+y+z+t)&1 and sites of one
    parity ordered by lexicographic index."""
    Lx, Ly, Lz, Lt = L
    t, z, y, x = np.meshgrid(np.arange(Lt), np.arange(Lz), np.arange(Ly), np.arange(Lx), indexing="ij")
    x, y, z, t = x.ravel(), y.ravel(), z.ravel(), t.ravel()  # lexicographic order, x fastest
    par = (x + y + z + t) & 1
    lex = np.arange(x.size)
    order = np.concatenate([lex[par == 0], lex[par == 1]])  # eo index -> lex index
    return order


def to_lex(field_eo, L):
    """[V4, ...] in even/odd order -> [Lt, Lz, Ly, Lx, ...]."""
    order = lex_coords(L)
    out = np.empty_like(field_eo)
    out[order] = field_eo
    return out.reshape((L[3], L[2], L[1], L[0]) + field_eo.shape[1:])


def to_eo(field_lex, L):
    order = lex_coords(L)
    flat = field_lex.reshape((-1,) + field_lex.shape[4:])
    return flat[order]


_AXIS = {0: 3, 1: 2, 2: 1, 3: 0}  # direction x,y,z,t -> array axis of [t,z,y,x]


def displace(v_eo, gauge_eo, direction, sign, L):
    """v_eo [V4,12] (c + 3 s), gauge_eo [4,V4,3,3].  plus: U_mu(x) v(x+mu); minus: U_mu(x-mu)^dag v(x-mu)."""
    v = to_lex(v_eo.reshape(-1, 4, 3), L)          # [t,z,y,x,s,c]
    U = to_lex(gauge_eo[direction], L)             # [t,z,y,x,r,c]
    ax = _AXIS[direction]
    if sign:
        out = np.einsum("...rc,...sc->...sr", U, np.roll(v, -1, axis=ax))
    else:
        Ub = np.roll(U, 1, axis=ax)
        out = np.einsum("...cr,...sc->...sr", Ub.conj(), np.roll(v, 1, axis=ax))
    return to_eo(out, L).reshape(-1, 12)


def contract(vL_eo, vR_eo, sigma):
    """[16, V4]: (1/sigma) vL^dag Gamma_G vR per site."""
    g = gamma_dense()
    l = vL_eo.reshape(-1, 4, 3)
    r = vR_eo.reshape(-1, 4, 3)
    return np.einsum("xbc,gba,xac->gx", l.conj(), g, r) / sigma


def compute_loop(evecs, sigma, gauge, entries, L):
    nLoop = 1 + sum(b - a + 1 for (_, _, a, b) in entries)
    V4 = evecs.shape[1]
    out = np.zeros((nLoop, 16, V4), dtype=complex)
    for n in range(evecs.shape[0]):
        out[0] += contract(evecs[n], evecs[n], sigma[n])
        iL = 1
        for (d, s, a, b) in entries:
            w = evecs[n]
            for k in range(1, b + 1):
                w = displace(w, gauge, d, s, L)
                if a <= k <= b:
                    out[iL + k - a] += contract(evecs[n], w, sigma[n])
            iL += b - a + 1
    return out


def momentum_projection(dataPos, mom, ftsign, L):
    """dataMom[im, G' + 16 iL, t] = sum_{xyz} sign[G] dataPos[iL, G, (xyz,t)] exp(i ftsign 2 pi p.x/L), with the
    reference's documented output map L(15-G) <- sign[G] T(G) (include/gamma.h:78-94)."""
    nLoop = dataPos.shape[0]
    minus = {3, 6, 9, 11, 12, 14}
    Lx, Ly, Lz, Lt = L
    lexd = to_lex(np.moveaxis(dataPos.reshape(nLoop * 16, -1), 0, 1), L)  # [t,z,y,x,idata]
    z, y, x = np.meshgrid(np.arange(Lz), np.arange(Ly), np.arange(Lx), indexing="ij")
    out = np.zeros((len(mom), 16 * nLoop, Lt), dtype=complex)
    for im, p in enumerate(mom):
        ph = np.exp(ftsign * 2j * np.pi * (p[0] * x / Lx + p[1] * y / Ly + p[2] * z / Lz))
        proj = np.einsum("tzyxd,zyx->dt", lexd, ph)
        for iL in range(nLoop):
            for G in range(16):
                out[im, (15 - G) + 16 * iL] = (-1 if G in minus else 1) * proj[G + 16 * iL]
    return out


def momentum_projection_mm(dataPos, mom, ftsign, L, loops=None):
    """Same result as momentum_projection for the loop slots `loops` (default: all), organised as one matrix product per
    time-slice so that the BASELINE-sized buffers (24^3x48: 33 loops, 33 momenta) take seconds, not minutes:
    out[im, G' + 16*k, t] for the k-th selected loop."""
    nLoop = dataPos.shape[0]
    loops = list(range(nLoop)) if loops is None else list(loops)
    minus = {3, 6, 9, 11, 12, 14}
    Lx, Ly, Lz, Lt = L
    z, y, x = np.meshgrid(np.arange(Lz), np.arange(Ly), np.arange(Lx), indexing="ij")
    ph = np.stack([np.exp(ftsign * 2j * np.pi * (p[0] * x / Lx + p[1] * y / Ly + p[2] * z / Lz)).ravel() for p in mom])  # [Nmom, V3]
    order = lex_coords(L)  # eo index -> lexicographic index
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)  # lexicographic index -> eo index
    inv = inv.reshape(Lt, Lx * Ly * Lz)
    out = np.zeros((len(mom), 16 * len(loops), Lt), dtype=complex)
    sign = np.array([-1.0 if G in minus else 1.0 for G in range(16)])
    for k, iL in enumerate(loops):
        for t in range(Lt):
            proj = ph @ dataPos[iL][:, inv[t]].T  # [Nmom, 16]
            out[:, 16 * k + 15 - np.arange(16), t] = proj * sign[None, :]
    return out
