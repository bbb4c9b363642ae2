python -m pytest tests -m gpu -q > gpurun_out/pytest_f.log 2>&1; tail -4 gpurun_out/pytest_f.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 300 gpurun_out/bench_f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_f.json')); print('VALUE', d['value'], 'E2E', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks']); 
for k,v in d['kernel_breakdown'].items(): print(k, v)"
