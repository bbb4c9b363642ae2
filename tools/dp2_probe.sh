# Two-GPU probe (gpurun --gpus N): NCCL data-parallel tests, the same box's N=1 and N=N bench values, what the gradient
# all-reduce costs (skip-all-reduce upper bound) and the device timeline of one data-parallel step.
TAG=${1:-r02}
N=${2:-2}
mkdir -p gpurun_out
if [ "${3:-tests}" = "tests" ]; then
timeout 900 python -m pytest tests/test_gpu_dp.py -q > gpurun_out/${TAG}_pytest_dp.log 2>&1
echo "pytest dp rc=$?"; tail -4 gpurun_out/${TAG}_pytest_dp.log | cut -c1-400
fi
FLAGS="--steps 60 --warmup 10 --no-configs --no-cpu-baseline --no-library-baseline --no-e2e"
run() {  # name, nproc, env...
  name=$1; np=$2; shift 2
  if [ "$np" = "1" ]; then
    env "$@" timeout 300 python bench.py --gpus 1 $FLAGS > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $np $FLAGS > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  fi
  python - <<PYEOF
import json
try:
    d = json.load(open('gpurun_out/${TAG}_${name}.json'))
    print('${name}', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'loss', d['loss'])
except Exception as e:
    print('${name} FAILED', e)
PYEOF
}
run n1 1 A=1
run n${N}_default $N A=1
run n${N}_skip_allreduce $N B200CD_DEBUG_SKIP_ALLREDUCE=1
run n${N}_no_tail_bucket $N B200CD_TAIL_FLUSH_DIV=0
env A=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29533 tools/dp_timeline.py dualstream > gpurun_out/${TAG}_dp_timeline_n${N}.json 2> gpurun_out/${TAG}_dp_timeline_n${N}.err
echo "timeline rc=$?"; head -c 1800 gpurun_out/${TAG}_dp_timeline_n${N}.json; echo
du -sh gpurun_out
