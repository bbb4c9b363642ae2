# same-box A/B of a tile table (in-process, alternating): all BASELINE configs, fast mode; dualstream precise
mkdir -p gpurun_out
T=${1:-profiles/r02_tile_table_v3.json}
: > gpurun_out/tiles_ab_v3.jsonl
for cfg in dualstream siamese dtsiamese mmcr dualstream; do
  timeout 400 python tools/step_ab.py $cfg fast ab --ab-table=$T 2>> gpurun_out/tiles_ab_v3.err | tee -a gpurun_out/tiles_ab_v3.jsonl | cut -c1-330
done
timeout 400 python tools/step_ab.py dualstream precise ab --ab-table=$T 2>> gpurun_out/tiles_ab_v3.err | tee -a gpurun_out/tiles_ab_v3.jsonl | cut -c1-330
