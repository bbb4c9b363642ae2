for t in 148 222 296 444; do
B200CD_WGRAD_CTAS=$t python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_e$t.json 2> gpurun_out/bench_e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_e$t.json')); print($t, 'VALUE', d['value'], 'ms', d['ms_per_step'], 'wgrad', d['kernel_breakdown']['wgrad'], d['kernel_breakdown']['wgrad_reduce'])"
done
