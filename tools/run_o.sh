python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "rc=$?"; tail -c 200 gpurun_out/bench_o.err; wc -l gpurun_out/bench_o.json; python -c "
import json; d=json.load(open('gpurun_out/bench_o.json')); print({k: d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks','cpu_baseline','vs_baseline','dtype')}); print(d['e2e']); print(d['roofline'])"
