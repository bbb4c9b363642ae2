python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1; tail -3 gpurun_out/pytest_c.log
timeout 300 python tools/gpu_probe.py --only pack_weights --perf > gpurun_out/probe_c.log 2>&1; grep -v '^{"pack' gpurun_out/probe_c.log | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): print(line); continue
    d=json.loads(line)
    for k,v in d.items():
        print(k, ' '.join(f'{a}={b:.0f}' for a,b in v.items() if isinstance(b,(int,float))))
"
