set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a1.log 2>&1; tail -5 gpurun_out/pytest_a1.log
B200CD_DUMP_CALLS=gpurun_out/calls_a1.txt python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a1.json 2> gpurun_out/bench_a1.err; tail -c 600 gpurun_out/bench_a1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_a1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step']); 
for k,v in d['kernel_breakdown'].items(): print(k, v)"
