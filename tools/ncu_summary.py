"""Condense an .ncu-rep (ncu --set full) into the few numbers the DESIGN / bench cite: duration, DRAM bytes, L2->SM
bytes, tensor-pipe activity, shared-memory data-pipe wavefronts by client (LSU loads / stores, tensor core), issue
activity, and the top stall reasons per warp role.  Usage: python tools/ncu_summary.py <report.ncu-rep> [> profiles/x.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print("== kernel:", m.get("Kernel Name", ("?", ""))[0][:200])
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_op_read_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
for w in want:
    for h in hdr:
        if h == w or h.endswith("." + w) or h.endswith(w):
            v, u = m[h]
            print(f"  {w} [{u}] = {v}")
            break
# stall reasons: sampled, aggregated over all instructions
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    sh = srows[1]
    ix = {h: i for i, h in enumerate(sh)}
    stall_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0.0 for h in stall_cols}
    for r in srows[2:]:
        for h in stall_cols:
            try:
                tot[h] += float(r[ix[h]])
            except (ValueError, IndexError):
                pass
    s = sum(tot.values()) or 1.0
    print("  warp stall samples (all warps):", ", ".join(f"{h[6:]} {100 * v / s:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
