set -u
O=gpurun_out
python bench.py > $O/r01_final_cornell.json 2> $O/r01_final_cornell.err || exit 1
python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_final_launches.csv \
    python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_wave --launch-skip 60 --launch-count 2 \
    -o $O/r01_final_cornell -f python bench.py --spp 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline > $O/ncu_full1.log 2>&1
tail -c 300 $O/r01_final_cornell.json
