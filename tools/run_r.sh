for cfg in "0 0" "1 0" "0 1" "0 0" "1 0" "0 1"; do set -- $cfg
B200CD_BN_FINE=$1 B200CD_BN_FINE_APPLY=$2 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_r.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); kb=d['kernel_breakdown']; print('FINE bwd=$1 apply=$2 VALUE', round(d['value'],1), round(d['ms_per_step'],3), 'bn_bwd', kb['bn_bwd']['ms'], 'bn_apply', kb['bn_apply']['ms'], 'bn_stats', kb['bn_stats']['ms'])"; done
