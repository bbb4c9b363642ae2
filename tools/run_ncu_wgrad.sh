tag=$1; shift
python tools/one_wgrad.py "$@" || exit 1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -c 3 -o gpurun_out/onew_$tag -f python tools/one_wgrad.py "$@" > gpurun_out/ncu_onew_$tag.log 2>&1
tail -2 gpurun_out/ncu_onew_$tag.log
ncu -i gpurun_out/onew_$tag.ncu-rep --page raw --csv > gpurun_out/onew_$tag.csv 2>/dev/null
python tools/ncu_key.py gpurun_out/onew_$tag.csv | grep -v "pcsamp\|smsp__average\|inst_executed_pipe_uniform\|sm__mem"
python tools/ncu_key.py gpurun_out/onew_$tag.csv | grep -i "sm__mem\|l1tex__data_pipe\|shared" | head -20
