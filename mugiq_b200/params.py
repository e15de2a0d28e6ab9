"""Loop-computation parameters: the Python mirror of MugiqLoopParam (/root/reference/include/mugiq.h:28-47),
the enums of include/enum_mugiq.h and the displacement / momentum parsing of the reference's driver
(tests/loop.cpp:607-746).  Error behaviour follows the reference: what is errorQuda there raises
MugiqError here, what is warningQuda there emits a Python warning and applies the same fix-up.
"""
import warnings
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

LOOP_FT_SIGN_MINUS, LOOP_FT_SIGN_PLUS = -1, 1                 # LoopFTSign, include/enum_mugiq.h:30-35
LOOP_CALC_TYPE_BLAS, LOOP_CALC_TYPE_OPT_KERNEL, LOOP_CALC_TYPE_BASIC_KERNEL = 0, 1, 2  # carried, never branched on
DISPLACE_TYPE_COVARIANT = 0                                    # DisplaceType, include/enum_mugiq.h:45-49
DISPLACE_FLAGS = ["+x", "-x", "+y", "-y", "+z", "-z", "+t", "-t"]  # DisplaceFlagArray, include/displace.h:21
GAMMA_NAMES = ["1", "g1", "g2", "g1g2", "g3", "g1g3", "g2g3", "g5g4",
               "g4", "g1g4", "g2g4", "g5g3", "g3g4", "g5g2", "g5g1", "g5"]  # GammaName, include/gamma.h:11-20
N_GAMMA = 16


class MugiqError(RuntimeError):
    """What the reference reports through errorQuda (fatal)."""


def which_displace(disp_str):
    """'+x'..'-t' -> (dir 0..3, sign 0=minus/1=plus).  Displace::WhichDisplaceFlag/Dir/Sign,
    lib/displace.cpp:137-202."""
    if disp_str not in DISPLACE_FLAGS:
        raise MugiqError(f"WhichDisplaceFlag: Cannot parse given displacement string = {disp_str}.")
    flag = DISPLACE_FLAGS.index(disp_str)
    return flag // 2, 1 - (flag % 2)


def parse_disp_entries(entry_string):
    """'+z:1,8;-x:3' -> (disp_entry, disp_str, disp_start, disp_stop) lists.  tests/loop.cpp:607-718."""
    if len(entry_string) == 0:
        raise MugiqError("Got option '--loop-do-nonlocal yes' but option --displace-entry-string is not set!")
    disp_entry, disp_str, disp_start, disp_stop = [], [], [], []
    for idx, ent in enumerate(entry_string.split(";")):
        split = ent.split(":")
        if len(split) != 2:
            raise MugiqError(f"Displacement entry {idx} has the Wrong format. Example of good entries: +z:1,8 , +x:3")
        try:
            lim = [int(s) for s in split[1].split(",")]
        except ValueError as exc:
            raise MugiqError(f"Wrong format of displacement entry {idx}. Example of good entries: +z:1,8 , +x:3") from exc
        if len(lim) == 0 or len(lim) > 2:
            raise MugiqError(f"Wrong format of displacement entry {idx}. Example of good entries: +z:1,8 , +x:3")
        disp_entry.append(ent)
        disp_str.append(split[0])
        disp_start.append(lim[0])
        disp_stop.append(lim[1] if len(lim) == 2 else lim[0])
    return disp_entry, disp_str, disp_start, disp_stop


def read_momenta(path):
    """Momenta text file, three ints per line.  tests/loop.cpp:723-746."""
    moms = []
    with open(path) as fh:
        for n, line in enumerate(fh):
            tok = line.split()
            try:
                moms.append([int(tok[0]), int(tok[1]), int(tok[2])])
            except (IndexError, ValueError) as exc:
                raise MugiqError(f"Incorrect file format in Line {n}") from exc
    return moms


def momenta_up_to(p2max):
    """All integer momenta with |p|^2 <= p2max in lexicographic order (SURVEY §8d synthetic inputs)."""
    r = int(p2max ** 0.5) + 1
    return [[px, py, pz] for px in range(-r, r + 1) for py in range(-r, r + 1) for pz in range(-r, r + 1)
            if px * px + py * py + pz * pz <= p2max]


@dataclass
class MugiqLoopParam:
    """include/mugiq.h:28-47.  `gauge` holds the four host link arrays (QDP order, one per direction)."""
    Nmom: int = 0
    momMatrix: List[List[int]] = field(default_factory=list)
    FTSign: int = LOOP_FT_SIGN_MINUS
    calcType: int = LOOP_CALC_TYPE_OPT_KERNEL
    writeMomSpaceHDF5: bool = False
    writePosSpaceHDF5: bool = False
    doMomProj: bool = False
    doNonLocal: bool = False
    disp_entry: List[str] = field(default_factory=list)
    disp_str: List[str] = field(default_factory=list)
    fname_mom_h5: str = ""
    fname_pos_h5: str = ""
    disp_start: List[int] = field(default_factory=list)
    disp_stop: List[int] = field(default_factory=list)
    gauge: Optional[Sequence] = None
    gauge_param: Optional[dict] = None

    def set_displacements(self, entry_string):
        self.disp_entry, self.disp_str, self.disp_start, self.disp_stop = parse_disp_entries(entry_string)
        self.doNonLocal = True

    def set_momenta(self, moms):
        self.momMatrix = [list(map(int, m)) for m in moms]
        self.Nmom = len(self.momMatrix)
        self.doMomProj = True


class LoopComputeParam:
    """Loop bookkeeping of Loop_Mugiq::LoopComputeParam (include/loop_mugiq.h:142-271)."""

    def __init__(self, prm: MugiqLoopParam, localL, comm_dim=(1, 1, 1, 1)):
        self.nG = N_GAMMA
        self.Nmom = prm.Nmom
        self.FTSign = prm.FTSign
        self.doMomProj = bool(prm.doMomProj)
        self.doNonLocal = bool(prm.doNonLocal)
        self.localL = [int(x) for x in localL]
        self.totalL = [self.localL[i] * int(comm_dim[i]) for i in range(4)]
        self.locT, self.totT = self.localL[3], self.totalL[3]
        self.locV3 = self.localL[0] * self.localL[1] * self.localL[2]
        self.totV3 = self.totalL[0] * self.totalL[1] * self.totalL[2]
        self.locV4 = self.locV3 * self.locT
        self.momMatrix = [p for m in prm.momMatrix for p in m] if self.doMomProj else []  # MOM_MATRIX_IDX order
        self.dispEntry, self.dispString, self.dispStart, self.dispStop = [], [], [], []
        self.nLoopPerEntry, self.nLoopOffset = [], []
        self.nLoop = 0
        if self.doNonLocal:
            self.nDispEntries = len(prm.disp_str)
            if self.nDispEntries != len(prm.disp_start) or self.nDispEntries != len(prm.disp_stop):
                raise MugiqError("Displacement string length not compatible with displacement limits length")
            for i in range(self.nDispEntries):
                start, stop = int(prm.disp_start[i]), int(prm.disp_stop[i])
                if start > stop:  # include/loop_mugiq.h:234-239
                    warnings.warn(f"Stop length is smaller than Start length for displacement {i}. Will switch lengths!")
                    start, stop = stop, start
                self.dispEntry.append(prm.disp_entry[i] if i < len(prm.disp_entry) else f"{prm.disp_str[i]}:{start},{stop}")
                self.dispString.append(prm.disp_str[i])
                self.dispStart.append(start)
                self.dispStop.append(stop)
                self.nLoopPerEntry.append(stop - start + 1)
                self.nLoopOffset.append(1 + sum(self.nLoopPerEntry[:i]))
                self.nLoop += self.nLoopPerEntry[i]
            self.nLoop += 1  # ultra-local
        else:
            self.nDispEntries = 0
            self.nLoop = 1
        self.nData = self.nLoop * self.nG

    def entries(self):
        """(dir, sign, start, stop) tuples for the C-ABI."""
        return [which_displace(s) + (a, b) for s, a, b in zip(self.dispString, self.dispStart, self.dispStop)]

    def loop_tags(self):
        """HDF5 group tag of every loop, in buffer order (lib/loop_mugiq.cpp:590-611, without the
        group2_tag[10] truncation)."""
        tags = ["disp_0"]
        for s, a, b in zip(self.dispString, self.dispStart, self.dispStop):
            tags += [f"disp_{s}_{k}" for k in range(a, b + 1)]
        return tags
