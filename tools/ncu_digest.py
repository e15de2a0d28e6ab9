"""Compact digest of an .ncu-rep: key raw metrics + top stall locations of the first kernel in the report.
usage: python tools/ncu_digest.py gpurun_out/prof.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for k in keys:
    if k in d:
        print(f"{k:80s} {d[k][1]} {d[k][0]}")
for k in hdr:
    if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued"):
        print(f"{k[33:]:30s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    n = int(r[ix["# Samples"]] or 0)
    br = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stalls), reverse=True)[:2]
    print(f"{r[ix['Address']][-5:]} {n:7d} {100*n/tot:5.1f}%  {br[0][1]}:{br[0][0]} {br[1][1]}:{br[1][0]} | {r[ix['Source']][:70]}")
