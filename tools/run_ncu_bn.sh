ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_dx|bn_bwd_reduce|bn_apply" -s 300 -c 12 -o gpurun_out/full_bn -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_bn.log 2>&1
tail -2 gpurun_out/ncu_full_bn.log | cut -c1-200
ncu -i gpurun_out/full_bn.ncu-rep --page raw --csv > gpurun_out/full_bn.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/full_bn.csv')))
hdr=rows[0]
want=["Kernel Name","Grid Size","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","lts__t_sector_hit_rate.pct","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct"]
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print(" | ".join((r[i][:60] if hdr[i]=="Kernel Name" else r[i]) for i in idx))
PY
