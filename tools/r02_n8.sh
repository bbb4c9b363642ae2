# 8-GPU check of the final data-parallel path on ONE box: N=1 and N=8 bench values, device timeline of one DP step.
N=${1:-8}
mkdir -p gpurun_out
FLAGS="--steps 60 --warmup 10 --no-configs --no-cpu-baseline --no-library-baseline"
timeout 200 python bench.py --gpus 1 $FLAGS --no-e2e > gpurun_out/r02_scale_same_box_n1.json 2> gpurun_out/r02_scale_same_box_n1.err; echo "n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N $FLAGS > gpurun_out/r02_scale_same_box_n$N.json 2> gpurun_out/r02_scale_same_box_n$N.err; echo "n$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/dp_timeline.py dualstream > gpurun_out/r02_dp_timeline_n$N.json 2> gpurun_out/r02_dp_timeline_n$N.err; echo "timeline rc=$?"
python - <<PYEOF
import json
for n in (1, $N):
    try:
        d = json.loads(open(f'gpurun_out/r02_scale_same_box_n{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', d.get('e2e', {}).get('value'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(n, 'FAILED', e)
PYEOF
head -c 1500 gpurun_out/r02_dp_timeline_n$N.json; echo
