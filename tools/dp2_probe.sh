# Two-GPU probe (gpurun --gpus 2): NCCL data-parallel tests, the same box's N=1 and N=2 bench values, what the gradient
# all-reduce costs (skip-all-reduce upper bound, NCCL CTA caps) and the device timeline of one data-parallel step.
TAG=${1:-r02}
N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py -q > gpurun_out/${TAG}_pytest_dp.log 2>&1
echo "pytest dp rc=$?"; tail -4 gpurun_out/${TAG}_pytest_dp.log | cut -c1-400
FLAGS="--steps 60 --warmup 10 --no-configs --no-cpu-baseline --no-library-baseline --no-e2e"
run() {  # name, nproc, env...
  name=$1; np=$2; shift 2
  if [ "$np" = "1" ]; then
    env "$@" timeout 300 python bench.py --gpus 1 $FLAGS > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $np $FLAGS > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  fi
  python - <<EOF
import json
try:
    d = json.load(open('gpurun_out/${TAG}_${name}.json'))
    print('${name}', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'loss', d['loss'])
except Exception as e:
    print('${name} FAILED', e)
EOF
}
run n1 1 A=1
run n${N}_default $N A=1
run n${N}_skip_allreduce $N B200CD_DEBUG_SKIP_ALLREDUCE=1
run n${N}_maxctas4 $N NCCL_MAX_CTAS=4
run n${N}_maxctas8 $N NCCL_MAX_CTAS=8
run n${N}_torchdist $N B200CD_NATIVE_COMM=0
for v in "default A=1" "maxctas4 NCCL_MAX_CTAS=4"; do
  set -- $v
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29533 tools/dp_timeline.py dualstream > gpurun_out/${TAG}_dp_timeline_n${N}_$1.json 2> gpurun_out/${TAG}_dp_timeline_n${N}_$1.err
  echo "timeline $1 rc=$?"; head -c 1500 gpurun_out/${TAG}_dp_timeline_n${N}_$1.json; echo
done
timeout 300 python tools/dp_timeline.py dualstream > gpurun_out/${TAG}_dp_timeline_n1.json 2> gpurun_out/${TAG}_dp_timeline_n1.err
head -c 600 gpurun_out/${TAG}_dp_timeline_n1.json; echo
du -sh gpurun_out
timeout 300 python tools/tiny_kernel_probe.py > gpurun_out/${TAG}_tiny_kernels.json 2> gpurun_out/${TAG}_tiny_kernels.err
cat gpurun_out/${TAG}_tiny_kernels.json; tail -3 gpurun_out/${TAG}_tiny_kernels.err
