"""One launch of each stage kernel at a representative size, for `ncu --set full` (north_star: HBM GB/s counters for
the contraction and displacement kernels, tensor-pipe utilisation for the projection).

    python tools/ncu_stage_kernels.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'tile_kernel|reorder|momproj_pos_dmma' \
        -o gpurun_out/r1_stage_kernels python tools/ncu_stage_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mugiq_b200 import ops  # noqa: E402
from mugiq_b200.params import momenta_up_to  # noqa: E402


def randc(*shape):
    return torch.randn(*shape, dtype=torch.complex128, device="cuda")


def main():
    L = (32, 32, 32, 64)
    V4 = L[0] * L[1] * L[2] * L[3]
    n = 8
    ev, ev2 = randc(n, V4, 12), randc(n, V4, 12)
    gauge = randc(4, V4, 3, 3)
    loop = torch.zeros(16, V4, dtype=torch.complex128, device="cuda")
    sig = [0.01 + 0.001 * i for i in range(n)]
    vl, vr = [ev[i] for i in range(n)], [ev2[i] for i in range(n)]
    ops.contract_batch(loop, vl, vr, sig, L, accumulate=False)   # contract_tile_kernel<double,false>
    ops.contract_batch(loop, vl, None, sig, L, accumulate=False)  # contract_tile_kernel<double,true>
    ops.displace_batch(vr, vl, gauge, 3, 1, L)                    # displace_tile_kernel<double>, +t
    ops.displace_batch(vr, vl, gauge, 0, 0, L)                    # -x
    torch.cuda.synchronize()
    del ev, ev2, gauge, loop, vl, vr

    L = (24, 24, 24, 48)
    V4 = L[0] * L[1] * L[2] * L[3]
    nLoop = 33
    mom = momenta_up_to(4)
    pos = randc(nLoop, 16, V4)
    ph = ops.phase_matrix_eo(mom, -1, L)
    out = ops.momproj_pos(pos, ph, nLoop, L)                      # momproj_pos_dmma_kernel<9>, config 3 shape
    mp = torch.empty(16 * nLoop * V4, dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(mp, pos[:9], 16 * 9, 9, L)               # reorder_mapgamma_kernel<double>
    torch.cuda.synchronize()
    print("ok", float(out.abs().sum()))


if __name__ == "__main__":
    main()
