# N-tile sweep of the CTA-pair kernel + same-box A/B of the resulting table + knob A/Bs (one GPU)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "forced_tiles or conv3x3 or bnbwd or convt" > $O/tiles_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/tiles_pytest.log
timeout 900 python tools/tile_sweep.py --out $O/tile_sweep.json --emit $O/tuned_tiles.json > $O/tile_sweep.log 2> $O/tile_sweep.err; echo "sweep rc=$?"
tail -2 $O/tile_sweep.log
: > $O/tiles_ab.jsonl
for cfg in dualstream siamese dtsiamese mmcr; do
  timeout 400 python tools/step_ab.py $cfg fast ab --ab-table=$O/tuned_tiles.json 2> $O/ab_$cfg.err | tee -a $O/tiles_ab.jsonl
done
for envs in "B200CD_FUSE_BN_BWD_MIN_PIXELS=8192" "B200CD_FUSE_BN_BWD=0" "B200CD_TAIL_FLUSH_DIV=0" "B200CD_TUNED_TILES=0"; do
  env $envs timeout 300 python tools/step_ab.py dualstream fast "$envs" 2>> $O/ab_knobs.err | tee -a $O/tiles_ab.jsonl
done
B200CD_TILE_TABLE=$O/tuned_tiles.json timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -k "step_parity and not precise" > $O/tiles_e2e.log 2>&1; echo "e2e rc=$?" | tee -a $O/tiles_e2e.log
tail -3 $O/tiles_e2e.log
