#!/bin/bash
# launch list (per-kernel durations) of the collision legs of the bench: sweep + edge batch (batch and resident form)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ig_|pq_|item_records|cover_|sphere_|compact_|scan_|sweep_|edge_check' -c 900 --csv \
    --log-file gpurun_out/r02_collide_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5 > gpurun_out/ncu_collide.log 2>&1
python scripts/launch_list.py gpurun_out/r02_collide_launches.csv 40
