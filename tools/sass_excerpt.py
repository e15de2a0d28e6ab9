"""SASS evidence of the built library: per kernel, the counts of the mnemonics that prove the Blackwell-native paths, and the
eigenvector loop of the fused kernel (opcode histogram + control-code stall counts).  usage: python tools/sass_excerpt.py > profiles/rN_sass_excerpt.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mugiq_b200", "lib", "libmugiq_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "DMMA", "DFMA", "DMUL", "FFMA", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
print("# SASS evidence, round 3 (cuobjdump -sass mugiq_b200/lib/libmugiq_b200.so; sm_100a, CUDA 12.9)")
print("#   UBLKCP = cp.async.bulk (linear bulk TMA)   UTMALDG = cp.async.bulk.tensor (tensor map)   SYNCS = mbarrier")
print("#   USETMAXREG = setmaxnreg (warp-specialised register split)   DMMA = mma.sync.m8n8k4.f64   LDL/STL = local memory (spills)\n")
funcs = re.split(r"\n\s+Function : ", txt)[1:]
for f, name in zip(funcs, names):
    ops = collections.Counter()
    for l in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", l)
        if m:
            t = m.group(1).split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += 1
    short = re.sub(r"mugiq_b200::", "", name)
    print(f"{short[:118]:118s} " + str({k: ops[k] for k in KEYS if ops[k]}))

# the eigenvector loop of loop_fused_kernel<double, 4, 0> (UL_ROT role: the loop BASELINE configs[1] runs)
sym = [s for s in re.findall(r"Function : (\S+)", txt) if "loop_fused_kernelIdLi4ELi0E" in s][0]
f = [x for x in funcs if x.startswith(sym)][0]
ins = []
lines = f.split("\n")
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        hi = int(m2.group(1), 16) if m2 else 0
        ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf))
        i += 2
        continue
    i += 1


def opof(s):
    t = s.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


loops = []
for a, s, st in ins:
    mb = re.search(r"BRA(\.U)?\s+(!?U?P\d,\s*)?0x([0-9a-f]+)", s)
    if mb and int(mb.group(3), 16) < a:
        body = [x for x in ins if int(mb.group(3), 16) <= x[0] <= a]
        nf = sum(1 for x in body if opof(x[1]) in ("DFMA", "DMUL"))
        if 380 <= nf <= 400:
            loops.append(body)
body = min(loops, key=len)
hist = collections.Counter(opof(s) for _, s, _ in body)
print(f"\n# eigenvector loop of loop_fused_kernel<double, 4, 0> (consumer_loop, ultra-local share rotated over the four roles):")
print(f"# {len(body)} instructions, sum of the control-code stall counts {sum(st for _, _, st in body)} cycles (lower bound for one warp alone)")
print("# " + str(sorted(hist.items(), key=lambda x: -x[1])))
print("# non-FP64 instructions of the loop, in order:")
for a, s, st in body:
    if opof(s) not in ("DFMA", "DMUL"):
        print(f"  /*{a:05x}*/ {s}")
