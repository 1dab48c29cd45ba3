#!/bin/bash
# quick perf line: tools/quick.sh <workload> <spp> [extra bench args]
W=$1; S=$2; shift 2
python bench.py --workload $W --spp $S --steps 2 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('%-12s %8.1f Mrays/s  %8.2f ms/step  trav %.3f ms/launch share %.2f  pairs/seg %.2f prims/seg %.2f' % ('$W', d['value'], d['ms_per_step'], r.get('mean_launch_ms',0), r.get('traverse_share_of_step',0), r.get('pairs_per_segment',0), r.get('prim_tests_per_segment',0)))"
