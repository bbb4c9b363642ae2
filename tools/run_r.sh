python -m pytest tests -m gpu -q -x -k "e2e or step or train or eval" > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
for s in 1 0 1 0; do
B200CD_WGRAD_SIDE_STREAM=$s python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('SIDE $s VALUE', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
done
tail -3 gpurun_out/bench_r.err
