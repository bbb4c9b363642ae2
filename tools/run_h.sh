n=$1
if [ "$n" = "1" ]; then
python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h$n.json 2> gpurun_out/bench_h$n.err
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 60 --warmup 5 > gpurun_out/bench_h$n.json 2> gpurun_out/bench_h$n.err
fi
tail -c 400 gpurun_out/bench_h$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h$n.json')); print('N', d['n_gpus'], 'VALUE', d['value'], 'E2E', d['e2e']['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['clocks'])"
