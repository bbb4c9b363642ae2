tag=$1
B200CD_DUMP_CALLS=gpurun_out/calls_$tag.txt python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('VALUE', d['value'], 'E2E', d['e2e']['value'], 'ms', d['ms_per_step'], d['clocks']); 
for k,v in d['kernel_breakdown'].items(): print(k, v)"
bash tools/run_ncu_list.sh $tag
