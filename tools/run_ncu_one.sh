# usage: bash tools/run_ncu_one.sh tag n H cin cout variant
tag=$1; shift
python tools/one_conv.py "$@" || exit 1
ncu --set full --clock-control none --import-source on -k regex:fprop -c 3 -o gpurun_out/one_$tag -f python tools/one_conv.py "$@" > gpurun_out/ncu_one_$tag.log 2>&1
tail -3 gpurun_out/ncu_one_$tag.log
ncu -i gpurun_out/one_$tag.ncu-rep --page raw --csv > gpurun_out/one_$tag.csv 2>/dev/null
python tools/ncu_key.py gpurun_out/one_$tag.csv
