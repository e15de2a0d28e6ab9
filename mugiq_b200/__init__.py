"""mugiq_b200 — B200-native (sm_100a) implementation of MuGiq's disconnected-loop hot path:
16-gamma eigenvector-sum contraction, covariant displacement, gamma/time-slice reorder, momentum projection.
The compute lives in mugiq_b200/lib/libmugiq_b200.so (hand-written CUDA behind the C-ABI of
include/mugiq_b200.h); PyTorch supplies device memory, streams and torch.distributed only."""
from . import _lib
from .lattice import Lattice
from .params import (MugiqLoopParam, LoopComputeParam, MugiqError, parse_disp_entries, which_displace, momenta_up_to,
                     GAMMA_NAMES, LOOP_FT_SIGN_MINUS, LOOP_FT_SIGN_PLUS, DISPLACE_TYPE_COVARIANT)

__all__ = ["_lib", "Lattice", "MugiqLoopParam", "LoopComputeParam", "MugiqError", "parse_disp_entries",
           "which_displace", "momenta_up_to", "GAMMA_NAMES", "LOOP_FT_SIGN_MINUS", "LOOP_FT_SIGN_PLUS",
           "DISPLACE_TYPE_COVARIANT"]
