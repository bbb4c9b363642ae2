python -m pytest tests -m gpu -q -x -k "e2e or step or train or eval or dropin or adamw" > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
for r in 0 1; do python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(\"value\", round(d[\"value\"],1), round(d[\"ms_per_step\"],3), \"e2e\", round(d[\"e2e\"][\"value\"],1), round(d[\"e2e\"][\"ms_per_step\"],3))"; done
python tools/host_probe.py dualstream 2>&1 | head -2
