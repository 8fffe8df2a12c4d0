mkdir -p gpurun_out
for g in 96 48 24 12 8; do echo "CSV_TILE_GRID=$g"; CSV_TILE_GRID=$g STRESS_TIME=1 timeout 300 python scripts/stress_shard.py 8 3 2>&1 | tail -1; done
for g in 128 64 32 16; do echo "CSV_WALK_GRID=$g"; CSV_WALK_GRID=$g STRESS_TIME=1 timeout 300 python scripts/stress_shard.py 8 3 2>&1 | tail -1; done
