i=0
for cfg in "4 0" "1 0" "2 0" "8 0" "4 1" "4 0"; do set -- $cfg; i=$((i+1))
B200CD_GRAD_BUCKETS=$1 B200CD_DEBUG_SKIP_ALLREDUCE=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/dp.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('buckets $1 skip $2: ', round(d['value'],1), round(d['ms_per_step'],3))"
done
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1: ', round(d['value'],1), round(d['ms_per_step'],3))"
