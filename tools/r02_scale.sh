#!/bin/bash
# r02 scaling pass on ONE 8-GPU box (gpurun --gpus 8): the bench line of every BASELINE.json config at 1 / 2 / 4 / 8 ranks
# (torchrun, NCCL merge; per_config block: C1, C3 full, C4 / C5 at the stated reduced spp), C4 at its full 1024 spp and C5 at
# 32 spp on 1 and 8 GPUs, the single-process product path rtw_render(gpus = 8), and the multi-GPU tests of the suite.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29510
run() {  # run <n> <tag> <bench args...>
  local n=$1 tag=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $O/r02_scale_${tag}_1.json 2> $O/r02_scale_${tag}_1.err
  else
    timeout 600 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > $O/r02_scale_${tag}_$n.json 2> $O/r02_scale_${tag}_$n.err
  fi
  echo "$tag N=$n rc=$? $(tail -c 300 $O/r02_scale_${tag}_$n.json | head -c 0)$(python - <<P
import json
try:
    d=json.loads([l for l in open('$O/r02_scale_${tag}_$n.json') if l.startswith('{')][-1])
    pc=d.get('per_config') or {}
    print('value %.0f e2e %.0f ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v.get('value', 0)) for k, v in pc.items()})
except Exception as e:
    print('unreadable', e)
P
)"
}
NS=${1:-"8 4 2 1"}; NF=${2:-"8 1"}   # rank counts for the default line / for the full-size C4 and C5 runs
for n in $NS; do run $n cornell --steps 3 --warmup 3 --no-cpu-baseline; done
for n in $NF; do run $n monument_full --workload monument --steps 2 --warmup 1 --no-cpu-baseline; done
for n in $NF; do run $n stress32 --workload stress --spp 32 --steps 2 --warmup 1 --no-cpu-baseline; done
# the product's own multi-GPU entry: one process, rtw_render(gpus = 8), peer stores into one frame
NG=$(nvidia-smi -L | wc -l)
timeout 300 python bench.py --gpus $NG --steps 3 --warmup 2 --no-cpu-baseline --no-per-config > $O/r02_scale_single_process_$NG.json 2> $O/r02_scale_single_process_$NG.err
echo "single process gpus=$NG rc=$?"; tail -c 400 $O/r02_scale_single_process_$NG.json
[ -z "${R02_SKIP_TESTS:-}" ] && timeout 300 python -m pytest tests -m gpu -q -k "gpus or clone or two" 2>&1 | tail -3 | tee $O/r02_scale_multigpu_tests.txt
