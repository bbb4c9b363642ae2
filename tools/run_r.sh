for s in 1 0 1 0; do
B200CD_BRANCH_STREAMS=$s python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('BRANCH $s VALUE', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
for s in 1 0; do
B200CD_BRANCH_STREAMS=$s python bench.py --config mmcr --batch 16 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mmcr16 BRANCH $s', round(d['value'],1))"
done
