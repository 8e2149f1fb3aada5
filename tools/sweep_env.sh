#!/bin/bash
# GPU box: A/B of run-time knobs of the CUDA module through environment variables.
# usage: tools/sweep_env.sh <workload> <spp> "VAR=a VAR2=b" "VAR=c" ...
WL=$1; SPP=$2; shift 2
for CFG in "$@"; do
  env $CFG python bench.py --workload $WL --spp $SPP --steps 2 --warmup 3 --e2e-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$CFG', '|', round(j['value'],1),'Mpaths/s', round(j['mrays_per_s'],1),'Mrays/s', {k[:8]:round(v['avg_launch_ms'],4) for k,v in j['kernels'].items()}, 'iters', j['wave_iterations'])
"
done
