set -x
mkdir -p gpurun_out
timeout 600 python scripts/diag_strong.py 2 wgs30x > gpurun_out/r2_diag_strong.txt 2>&1; tail -40 gpurun_out/r2_diag_strong.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or adversarial or prepass or refuses or multi_contig or pileups or empty or sharding or long_reads" 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-map > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; tail -c 1800 gpurun_out/r2_b3.json; tail -5 gpurun_out/r2_b3.err
for g in 96 8 7 6; do CSV_SIDE_PRIO=0 CSV_TILE_GRID=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b3_g$g.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_b3_g$g.json'));print('prio0 grid $g', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"; done
for g in 8 7; do CSV_TILE_GRID=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r2_b3_h$g.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_b3_h$g.json'));print('prioHI grid $g', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"; done
