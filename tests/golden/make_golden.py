"""Regenerates tests/golden/loop_4x4x4x8.npz: seeds + outputs of the 4^3x8 / 16-eigenvector configuration
(BASELINE.json configs[0]) on which the C++ oracle and the independent numpy restatement agree to 1e-13.
The reference has no golden vectors of its own and cannot be built here (QUDA absent), so this fixture pins
the oracle against regressions; the fixture that pins it to the reference's own kernels is ref_kernels_4x4x4x8.npz
(make_ref_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mugiq_b200 import synth  # noqa: E402
from mugiq_b200.params import momenta_up_to  # noqa: E402
from oracle import numpy_check as npc  # noqa: E402
from oracle import oracle as orc  # noqa: E402

L = (4, 4, 4, 8)
NEV = 16
SEED = 2026
ENTRIES = [(0, 1, 1, 1), (0, 0, 1, 1), (1, 1, 1, 1), (1, 0, 1, 1), (2, 1, 1, 1), (2, 0, 1, 1), (3, 1, 1, 2), (3, 0, 1, 2)]

ev = synth.random_evecs_np(L, NEV, seed=SEED)
sig = synth.sigmas(NEV)
U = synth.random_gauge(L, seed=SEED)
a = orc.compute_loop(ev, sig, U, ENTRIES, L)
b = npc.compute_loop(ev, sig, U, ENTRIES, L)
err = np.abs(a - b).max() / np.abs(b).max()
assert err < 1e-13, err
mom = momenta_up_to(2)
dm = npc.momentum_projection(a, mom, -1, L)
# keep the fixture small: full momentum-space result + a strided sample and per-(loop,gamma) sums of dataPos
np.savez_compressed(os.path.join(os.path.dirname(__file__), "loop_4x4x4x8.npz"),
                    L=np.array(L), nEv=NEV, seed=SEED, entries=np.array(ENTRIES), mom=np.array(mom), ftsign=-1,
                    dataPos_sample=a[:, :, ::37], dataPos_sums=a.sum(axis=2), dataMom=dm,
                    ev_checksum=ev.sum(), gauge_checksum=U.sum())
print("wrote golden fixture, oracle-vs-numpy rel err", err)
