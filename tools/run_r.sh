python -m pytest tests -m gpu -q -x -k "pair or conv or e2e or step or bn" > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
python tools/gpu_probe.py --perf --fprop-only --only none 2>&1 | grep fprop_pair | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    for k,v in d.items(): print(k, round(v['fprop_pair_tflops']), round(v['fprop_pair_us'],1))
"
for i in 1 2; do python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('VALUE', d['value'], d['ms_per_step']); [print(k,v) for k,v in d['kernel_breakdown'].items() if 'fprop' in k or 'dgrad' in k or 'gemm1' in k]"; done
tail -3 gpurun_out/bench_r.err
