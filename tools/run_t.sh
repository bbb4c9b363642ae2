python -m pytest tests -m gpu -q -x -k "pack or e2e" > gpurun_out/pytest_t.log 2>&1; tail -3 gpurun_out/pytest_t.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('VALUE', round(d['value'],1), round(d['ms_per_step'],3)); [print(k,v) for k,v in d['kernel_breakdown'].items() if 'pack' in k]"
