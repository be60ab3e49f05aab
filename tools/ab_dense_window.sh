#!/bin/bash
# DRAM traffic + time of the dense pair kernel (split precision, bench shape) for several walker windows
for w in "$@"; do
  echo "== window $w"
  R4D_DENSE_WALKER_WINDOW=$w ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dense2_kernel -s 1 -c 1 python bench.py --scorers dense --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-aux 2>&1 | grep "dram__bytes_read\|gpu__time\|hit_rate"
done
