# One-GPU evidence pass (run under gpurun from the repo root): GPU tests, the default bench line, the ncu launch list of
# the same bench command, per-launch DRAM traffic of one training step, one `--set full` capture of the dominant kernel.
# Outputs land in gpurun_out/ (scratch); the summaries are copied / joined into profiles/ afterwards.
# gpurun copies back at most 64 MiB: the .ncu-rep files are exported to CSV on the box and only the small one is kept.
TAG=${1:-r02}
mkdir -p gpurun_out
if [ "${2:-all}" != "nopytest" ]; then
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
fi
timeout 600 python tools/sanitize_kernels.py --guard > gpurun_out/${TAG}_redzone.log 2>&1
echo "red-zone rc=$?"; tail -3 gpurun_out/${TAG}_redzone.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/${TAG}_bench_dualstream_16.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench_dualstream_16.json; echo
# launch list of the same command (short run; a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
  --log-file gpurun_out/${TAG}_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-library-baseline --no-e2e \
  > gpurun_out/${TAG}_ncu_launches_bench.out 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/${TAG}_ncu_launches_bench.csv
# per-launch DRAM traffic of one eager step on one stream
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -f -o gpurun_out/${TAG}_step_all python tools/ncu_step_traffic.py run dualstream > gpurun_out/${TAG}_step_all.out 2>&1
echo "ncu traffic rc=$?"; tail -2 gpurun_out/${TAG}_step_all.out | cut -c1-300
ncu -i gpurun_out/${TAG}_step_all.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_all_raw.csv 2> gpurun_out/${TAG}_step_all_raw.err
rm -f gpurun_out/${TAG}_step_all.ncu-rep
# full capture of the dominant kernel: 6 launches out of the middle of the step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fprop_pair_kernel -s 10 -c 6 \
  -f -o gpurun_out/${TAG}_full_pair python tools/ncu_step_traffic.py run dualstream > gpurun_out/${TAG}_full_pair.out 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
ncu -i gpurun_out/${TAG}_full_pair.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_pair_raw.csv 2>/dev/null
du -sh gpurun_out
