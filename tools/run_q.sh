python -m pytest tests -m gpu -q > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_q_dualstream_16.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q_dualstream_16.json')); print('dualstream B=16', 'VALUE', round(d['value'],1), 'E2E', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'step_roofline', round(d['step_roofline']['frac_of_tensor_peak'],3), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3), d['clocks'], d['cpu_baseline']['value'])"
for c in "siamese 32" "dtsiamese 8" "dtsiamese_ssl 8" "mmcr 64" "siamese 8"; do set -- $c
python bench.py --config $1 --batch $2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q_$1_$2.json 2> gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q_$1_$2.json')); print('$1 B=$2', 'VALUE', round(d['value'],1), 'E2E', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'step_roofline', round(d['step_roofline']['frac_of_tensor_peak'],3), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done
