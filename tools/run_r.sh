python -m pytest tests -m gpu -q -x -k "wgrad or e2e or step" > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
python tools/one_wgrad.py 32 256 64 64 5; B200CD_WGRAD_MSTACK=0 python tools/one_wgrad.py 32 256 64 64 5
for s in 1 0; do
B200CD_WGRAD_MSTACK=$s python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MSTACK $s VALUE', d['value'], d['ms_per_step']); [print(k,v) for k,v in d['kernel_breakdown'].items() if 'wgrad' in k]"
done
tail -3 gpurun_out/bench_r.err
