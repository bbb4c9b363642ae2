# graph-vs-eager matrix on one GPU
mkdir -p gpurun_out
for cfg in "$@"; do
i=0
for envs in "B200CD_WGRAD_SIDE_PER_BRANCH=0" "B200CD_WGRAD_SIDE_PER_BRANCH=1" "B200CD_WGRAD_SIDE_PER_BRANCH=1 B200CD_BRANCH_STREAMS=2"; do
  env $envs timeout 300 python tools/graph_vs_eager.py $cfg 2> gpurun_out/gve_$i.err | tee -a gpurun_out/gve3_$cfg.jsonl
  i=$((i+1))
done
done
