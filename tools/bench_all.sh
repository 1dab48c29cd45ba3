#!/bin/bash
# Full r01 measurement pass (run on the GPU box): five BASELINE.json configs, then the ncu launch list and one
# `--set full` capture of the same bench command (only after the plain run exited 0).
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/r01_final_cornell.json 2> $O/r01_final_cornell.err || exit 1
python bench.py --workload cow > $O/r01_final_cow.json 2> $O/r01_final_cow.err
python bench.py --workload jumpy-balls > $O/r01_final_jumpy.json 2> $O/r01_final_jumpy.err
python bench.py --workload monument --spp 32 --steps 2 > $O/r01_final_monument.json 2> $O/r01_final_monument.err
python bench.py --workload stress --spp 8 --steps 2 > $O/r01_final_stress.json 2> $O/r01_final_stress.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r01_final_reference.json 2> $O/r01_final_reference.err
# ncu: launch list (gpu__time_duration) of a short run of the default bench command
python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_final_launches.csv \
    python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_wave --launch-skip 60 --launch-count 2 \
    -o $O/r01_final_cornell -f python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_wave --launch-skip 16 --launch-count 2 \
    -o $O/r01_final_stress -f python bench.py --workload stress --spp 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_wave --launch-skip 30 --launch-count 2 \
    -o $O/r01_final_cow -f python bench.py --workload cow --spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_full3.log 2>&1
for f in cornell cow jumpy monument stress reference; do tail -c 400 $O/r01_final_$f.json; echo; done
