python -m pytest tests -m gpu -q -x > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_k.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('VALUE', d['value'], d['ms_per_step']); [print(k,v) for k,v in d['kernel_breakdown'].items()]"
tail -3 gpurun_out/bench_k.err
