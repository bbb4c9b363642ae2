# quick GPU regression subset: graph-replay step tests, DP on one GPU, smoke
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -x -k "fused_graph or siamese_dropin or baseline_size" > gpurun_out/quick_e2e.log 2>&1; echo "e2e rc=$?"; tail -3 gpurun_out/quick_e2e.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_dp.py -q -x > gpurun_out/quick_dp.log 2>&1; echo "dp rc=$?"; tail -3 gpurun_out/quick_dp.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
