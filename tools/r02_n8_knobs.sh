# N-GPU A/B of data-parallel knobs on ONE box (short runs; charged N x box time)
N=${1:-8}
mkdir -p gpurun_out
FLAGS="--steps 40 --warmup 8 --no-configs --no-cpu-baseline --no-library-baseline --no-e2e"
: > gpurun_out/r02_n${N}_knobs.jsonl
i=0
for envs in "A=1" "NCCL_MAX_CTAS=8" "B200CD_GRAD_BUCKETS=2"; do
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29511+i)) \
    bench.py --gpus $N $FLAGS > gpurun_out/knob_$i.json 2> gpurun_out/knob_$i.err
  python - <<PYEOF
import json
try:
    d = json.loads(open('gpurun_out/knob_$i.json').read().strip().splitlines()[-1])
    r = {"env": "$envs", "n_gpus": d["n_gpus"], "value": d["value"], "ms_per_step": d["ms_per_step"], "clocks": d.get("clocks")}
except Exception as e:
    r = {"env": "$envs", "failed": str(e)}
print(json.dumps(r))
open('gpurun_out/r02_n${N}_knobs.jsonl', 'a').write(json.dumps(r) + "\n")
PYEOF
  i=$((i+1))
done
