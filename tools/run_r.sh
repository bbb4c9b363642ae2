python -m pytest tests -m gpu -q -x -k "e2e or step or train or eval or dp" > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
for cfg in "dtsiamese 8" "dtsiamese_ssl 8"; do set -- $cfg
for s in 1 0 1 0; do
B200CD_BRANCH_STREAMS=$s python bench.py --config $1 --batch $2 --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 BRANCH $s VALUE', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done; done
tail -3 gpurun_out/bench_r.err
