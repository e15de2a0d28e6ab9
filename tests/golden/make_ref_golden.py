"""Generates tests/golden/ref_kernels_4x4x4x8.npz ON A GPU BOX: outputs of the reference's OWN CUDA kernels and wrappers
(oracle/_ref/libmugiq_ref.so = /root/reference/lib/{contract_wrappers,mugiq_contract_kernels,mugiq_displace_kernels,
mugiq_util_kernels}.cu compiled unmodified against oracle/quda_shim) on seeded inputs of BASELINE.json configs[0]
(4^3x8 random SU(3) lattice, 16 random eigenvectors).  The CPU suite (tests/test_golden_ref.py) checks the oracle
against these vectors, which pins it to the reference's kernel code; what stays an assumption is QUDA's side of the
accessors (oracle/quda_shim/quda_shim_core.h).

    gpurun -- python tests/golden/make_ref_golden.py      # writes gpurun_out/ref_kernels_4x4x4x8.npz; copy it here
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mugiq_b200 import synth  # noqa: E402
from mugiq_b200.params import momenta_up_to  # noqa: E402
from oracle import ref_kernels as ref  # noqa: E402

L = (4, 4, 4, 8)
NEV = 16
SEED = 2026
# config 0 is ultra-local; the displaced entries pin the displacement kernel too (lengths 1..2, both signs)
ENTRIES = [(0, 1, 1, 1), (0, 0, 1, 1), (1, 1, 1, 1), (1, 0, 1, 1), (2, 1, 1, 1), (2, 0, 1, 1), (3, 1, 1, 2), (3, 0, 1, 2)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gauge_dev(U):
    # host QDP order [mu][parity][x_cb][3][3] is what the shim's gauge accessor indexes: a plain copy, no kernel involved
    return dev(np.stack([U[mu] for mu in range(4)]))


ev = synth.random_evecs_np(L, NEV, seed=SEED)
sig = synth.sigmas(NEV)
U = synth.random_gauge(L, seed=SEED)
gd = gauge_dev(U)
out = {}
for order in (2, 4):
    q = [ref.site_to_quda(dev(ev[i]), order) for i in range(NEV)]
    pos = ref.compute_loop(q, sig, gd, ENTRIES, L, order)
    out[f"dataPos_order{order}"] = pos.cpu().numpy()
assert np.array_equal(out["dataPos_order2"], out["dataPos_order4"]), "FLOAT2 and FLOAT4 runs of the reference differ"
pos_h = out["dataPos_order2"]
nLoop, V4, V3, Lt = pos_h.shape[0], pos_h.shape[2], L[0] * L[1] * L[2], L[3]

# single kernels
q0 = ref.site_to_quda(dev(ev[0]), 2)
disp = np.zeros((4, 2, V4, 12), dtype=np.complex128)
for d in range(4):
    for s in (0, 1):
        o = torch.zeros_like(q0)
        ref.displace(o, q0, gd, d, s, L, 2, True)
        o2 = torch.zeros_like(q0)
        ref.displace(o2, q0, gd, d, s, L, 2, False)
        assert torch.equal(o, o2), "extended and non-extended gauge branches of the reference kernel differ"
        disp[d, s] = ref.quda_to_site(o, 2).cpu().numpy()
mp = torch.zeros(V3 * 16 * nLoop * Lt, dtype=torch.complex128, device="cuda")
ref.reorder_mapgamma(mp, dev(pos_h), nLoop, L)
mom = momenta_up_to(2)
ph = {sgn: ref.phase_matrix(mom, sgn, L).cpu().numpy() for sgn in (-1, 1)}
dm = torch.matmul(dev(ph[-1]), mp.reshape(V3, 16 * nLoop * Lt)).cpu().numpy().reshape(len(mom), 16 * nLoop, Lt)

dst = os.path.join(ROOT, "gpurun_out", "ref_kernels_4x4x4x8.npz")
os.makedirs(os.path.dirname(dst), exist_ok=True)
mp_h = mp.cpu().numpy()
np.savez_compressed(dst, L=np.array(L), nEv=NEV, seed=SEED, entries=np.array(ENTRIES), mom=np.array(mom),
                    dataPos_sample=pos_h[:, :, ::5], dataPos_sums=pos_h.sum(axis=2),
                    displace_sample=disp[:, :, ::3], displace_sums=disp.sum(axis=2),
                    reorder_sample=mp_h[::101], reorder_weighted_sum=(mp_h * np.arange(1, mp_h.size + 1)).sum(),
                    phase_minus=ph[-1], phase_plus=ph[1], dataMom=dm,
                    ev_checksum=ev.sum(), gauge_checksum=U.sum(),
                    device=torch.cuda.get_device_name(0))
print("wrote", dst, "dataPos", pos_h.shape, "max|dataPos|", np.abs(pos_h).max())
