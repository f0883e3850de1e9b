#!/bin/bash
# launch list (per-kernel durations) of the collision legs of the bench: sweep + edge batch (batch and resident form)
set -e
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5 > gpurun_out/ncu_collide_plain.json 2> gpurun_out/ncu_collide_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pq_|item_records|cover_|sphere_|compact_|scan_|sweep_|edge_' -c 600 --csv \
    --log-file gpurun_out/r02_collide_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5 > gpurun_out/ncu_collide.log 2>&1
python scripts/launch_list.py gpurun_out/r02_collide_launches.csv 60
