# ncu evidence for bench.py: launch list (durations) + one full capture of the dominant kernel. usage: bash tools/run_profiles.sh <tag>
tag=$1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_$tag.log 2>&1
tail -2 gpurun_out/ncu_bench_$tag.log | cut -c1-300
python tools/agg_ncu.py gpurun_out/launches_bench_$tag.csv 8 | head -40
ncu --set full --clock-control none --import-source on -k regex:fprop_pair -s 60 -c 6 -o gpurun_out/full_pair_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-300
ncu -i gpurun_out/full_pair_$tag.ncu-rep --page raw --csv > gpurun_out/full_pair_$tag.csv 2>/dev/null
for r in 2 3 4 5 6 7; do python tools/ncu_key.py gpurun_out/full_pair_$tag.csv $r | grep -E "Kernel Name|Grid Size|gpu__time_duration|dram__bytes|xbar2l1tex_read|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|lts__t_sector_hit"; echo; done
