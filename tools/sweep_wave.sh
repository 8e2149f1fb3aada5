#!/bin/bash
# GPU box: sweep the wavefront capacity (path slots in flight). Smaller waves keep the path state L2-resident.
WL=${1:-c2}; SPP=${2:-256}
for P in 262144 524288 1048576 2097152 4194304; do
  python bench.py --workload $WL --spp $SPP --steps 2 --warmup 3 --e2e-steps 1 --no-cpu --wave-capacity $P 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('P=$P', round(j['value'],1),'Mpaths/s', round(j['mrays_per_s'],1),'Mrays/s', {k[:8]:round(v['avg_launch_ms'],4) for k,v in j['kernels'].items()}, 'iters', j['wave_iterations'])
"
done
