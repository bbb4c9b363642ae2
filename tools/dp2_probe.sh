set -x
timeout 900 python -m pytest tests/test_gpu_dp.py -q -x > gpurun_out/pytest_dp2_r02.log 2>&1; tail -15 gpurun_out/pytest_dp2_r02.log | cut -c1-600
for combo in "1 1" "0 1" "1 0" "0 0"; do
  set -- $combo
  B200CD_DP_GRAPH=$1 B200CD_NATIVE_COMM=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --no-configs > gpurun_out/bench_n2_graph$1_native$2.json 2> gpurun_out/bench_n2_graph$1_native$2.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/bench_n2_graph$1_native$2.json')); print('graph$1 native$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])
except Exception as e: print('graph$1 native$2 FAILED', e)
"
  tail -3 gpurun_out/bench_n2_graph$1_native$2.err | cut -c1-300
done
