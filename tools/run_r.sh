for i in 1 2 3; do
python bench.py --config dtsiamese --batch 8 --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dtsiamese VALUE', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))"
done
python bench.py --config dtsiamese --batch 8 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dtsiamese 30 steps VALUE', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))"
python tools/host_probe.py dtsiamese 2>&1 | head -1
