for s in 1 0 1 0; do
B200CD_WGRAD_SIDE_STREAM=$s python bench.py --config siamese --batch 8 --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('siamese8 SIDE $s VALUE', round(d['value'],1), round(d['ms_per_step'],3))"
done
for s in 1 0; do
B200CD_WGRAD_SIDE_STREAM=$s python bench.py --config dtsiamese --batch 8 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_t.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dtsiamese8 SIDE $s VALUE', round(d['value'],1), round(d['ms_per_step'],3))"
done
