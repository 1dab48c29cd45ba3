#!/bin/bash
# r02 ncu pass (run on the GPU box, one GPU): launch list of the default bench command, then one `--set full` capture of
# the dominant kernels of each config.  Every ncu command follows a plain run of the same command that exited 0.
# usage: tools/r02_ncu.sh <tag>     -> gpurun_out/<tag>_*.ncu-rep, gpurun_out/<tag>_launches.csv
set -u
T=${1:-r02}
O=gpurun_out
mkdir -p $O
Q="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline --no-per-config"
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_plain.json 2> $O/${T}_plain.err || { tail -5 $O/${T}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_launches.log 2>&1
python bench.py --spp 100 $Q > /dev/null 2>&1 && \
$NCU --kernel-name regex:k_mega -c 1 -o $O/${T}_cornell python bench.py --spp 100 $Q > $O/${T}_ncu1.log 2>&1
python bench.py --workload cow --spp 64 $Q > /dev/null 2>&1 && \
$NCU --kernel-name regex:k_wave --launch-skip 20 --launch-count 2 -o $O/${T}_cow python bench.py --workload cow --spp 64 $Q > $O/${T}_ncu2.log 2>&1
python bench.py --workload monument --spp 16 $Q > /dev/null 2>&1 && \
$NCU --kernel-name regex:k_wave --launch-skip 20 --launch-count 2 -o $O/${T}_monument python bench.py --workload monument --spp 16 $Q > $O/${T}_ncu3.log 2>&1
python bench.py --workload stress --spp 4 $Q > /dev/null 2>&1 && \
$NCU --kernel-name regex:k_wave --launch-skip 20 --launch-count 2 -o $O/${T}_stress python bench.py --workload stress --spp 4 $Q > $O/${T}_ncu4.log 2>&1
ls -la $O/${T}_*
