python -m pytest tests -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; tail -3 gpurun_out/pytest_i.log
for pdl in 1 0; do
B200CD_PDL=$pdl python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_i$pdl.json 2> gpurun_out/bench_i$pdl.err; tail -c 300 gpurun_out/bench_i$pdl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_i$pdl.json')); print('PDL', $pdl, 'VALUE', d['value'], 'E2E', d['e2e']['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'])"
done
