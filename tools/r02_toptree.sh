#!/bin/bash
# r02: the shared-memory top of the tree (north_star item, RTW_TOP_TREE) measured against the shipped build on the same box:
# A/B timings, the parity suite through the macro path, and one `ncu --set full` launch of the traversal kernel each.
# Needs lib/ab_top64 (tools/ab_build.sh top64 -DRTW_TOP_TREE=64) and lib/ab_top16.
set -u
O=gpurun_out
L=raytracer-weekend_b200/lib
Q="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline --no-per-config"
python tools/ab.py --variants base top16 top64 --work cow:64 monument:16 jumpy-balls:100 > $O/r02_toptree_ab.txt 2>&1
cat $O/r02_toptree_ab.txt
for w in cow monument; do
  spp=64; [ $w = monument ] && spp=16
  ncu --set full --clock-control none --import-source on -f --kernel-name regex:k_wave_traverse --launch-skip 10 --launch-count 1 \
      -o $O/r02_top_base_$w python bench.py --workload $w --spp $spp $Q > $O/r02_top_ncu_base_$w.log 2>&1
done
cp $L/librtw_cuda.so $L/librtw_cuda.so.base
cp $L/ab_top64/librtw_cuda.so $L/librtw_cuda.so
python -m pytest tests -m gpu -x -q -k "not console_app" 2>&1 | tail -3 | tee $O/r02_toptree_parity.txt
for w in cow monument; do
  spp=64; [ $w = monument ] && spp=16
  ncu --set full --clock-control none --import-source on -f --kernel-name regex:k_wave_traverse --launch-skip 10 --launch-count 1 \
      -o $O/r02_top_top64_$w python bench.py --workload $w --spp $spp $Q > $O/r02_top_ncu_top64_$w.log 2>&1
done
cp $L/librtw_cuda.so.base $L/librtw_cuda.so
