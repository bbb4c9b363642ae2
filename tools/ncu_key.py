"""Print the metrics that matter from an `ncu --page raw --csv` export (last launch). usage: python tools/ncu_key.py file.csv [row]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
r = rows[int(sys.argv[2])] if len(sys.argv) > 2 else rows[-1]
want = re.compile(
    r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|sm__cycles_elapsed.avg$|sm__cycles_active.avg$|"
    r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|sm__throughput.avg.pct|"
    r"l1tex__m_xbar2l1tex_read_bytes.sum$|l1tex__m_l1tex2xbar_write_bytes.sum$|lts__t_bytes.sum$|lts__t_sector_hit_rate.pct|"
    r"lts__throughput.avg.pct|dram__bytes_read.sum$|dram__bytes_write.sum$|dram__throughput.avg.pct|"
    r"launch__cluster|launch__grid_size|launch__occupancy_cluster|launch__registers|launch__shared_mem_per_block_dynamic|"
    r"smsp__average_warp.*_per_issue_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$|sm__warps_active.avg.pct|"
    r"smsp__inst_executed.sum$|sm__ctas_launched|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__warp_issue_stalled.*|"
    r"sm__inst_executed_pipe_uniform|sm__sass_inst_executed_op_shared.*sum$|sm__mem.*|smsp__pcsamp_warps_issue_stalled_[a-z_]+$)")
for i, h in enumerate(hdr):
    hh = h.split(".TriageCompute.")[-1]
    if want.search(hh) and r[i] not in ("", "0"):
        print(f"{hh[:90]:90s} {units[i]:10s} {r[i]}")
