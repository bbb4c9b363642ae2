python -m pytest tests -m gpu -q > gpurun_out/pytest_g.log 2>&1; tail -3 gpurun_out/pytest_g.log
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 300 gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print('VALUE', d['value'], 'E2E', d['e2e'], 'ms', d['ms_per_step'], d['clocks'], d['cpu_baseline'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
