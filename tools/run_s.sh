# full refresh: GPU tests, smoke, benches of all configs, ncu launch list + full capture of the pair kernel
bash tools/run_q.sh
bash tools/run_profiles.sh v4 > gpurun_out/profiles_v4.log 2>&1; tail -60 gpurun_out/profiles_v4.log
