# launch list of two eager steps (per-kernel GPU durations); usage: bash tools/run_ncu_list.sh <tag> [config] [batch]
tag=${1:-x}; cfg=${2:-dualstream}; b=${3:-}
python tools/ncu_step.py $cfg $b > gpurun_out/ncu_plain_$tag.log 2>&1 || { tail -5 gpurun_out/ncu_plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv python tools/ncu_step.py $cfg $b > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
python tools/agg_ncu.py gpurun_out/launches_$tag.csv 2
