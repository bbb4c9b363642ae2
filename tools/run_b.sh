timeout 300 python tools/gpu_probe.py --only "$1" $2 $3 > gpurun_out/probe_b.log 2>&1; echo rc=$?; tail -60 gpurun_out/probe_b.log | cut -c1-700
