#!/bin/bash
# r02 final measurement pass (one GPU): the default bench line (what the driver records), every config as a line of its own
# with its CPU arm, the reference arm, and the ptxas / SASS tables of the build that produced them.
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/r02_final_bench_cornell.json 2> $O/r02_final_bench_cornell.err || { tail -5 $O/r02_final_bench_cornell.err; exit 1; }
python bench.py --workload cow > $O/r02_final_bench_cow.json 2> $O/r02_final_bench_cow.err
python bench.py --workload jumpy-balls > $O/r02_final_bench_jumpy.json 2> $O/r02_final_bench_jumpy.err
python bench.py --workload monument --spp 32 --steps 2 > $O/r02_final_bench_monument.json 2> $O/r02_final_bench_monument.err
python bench.py --workload stress --spp 4 --steps 2 > $O/r02_final_bench_stress.json 2> $O/r02_final_bench_stress.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_bench_reference.json 2> $O/r02_final_bench_reference.err
for f in cornell cow jumpy monument stress reference; do
  python - <<P
import json
try:
    d=json.loads([l for l in open('$O/r02_final_bench_$f.json') if l.startswith('{')][-1])
    r=d.get('roofline') or {}; c=d.get('cpu_baseline') or {}
    print('$f', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms %.1f' % d['ms_per_step'], 'cpu %.1f (%s cores)' % (c.get('value', 0), c.get('cores')),
          'roofline', r.get('kernel'), 'share %.2f' % (r.get('share_of_step') or 0), 'frac %.3f' % (r.get('frac') or 0), 'traffic', r.get('traffic'))
except Exception as e:
    print('$f unreadable', e)
P
done
