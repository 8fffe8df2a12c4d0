set -x
mkdir -p gpurun_out
bash scripts/r2_run_strong.sh 2
cp gpurun_out/r2_strong_2.json gpurun_out/r2_strong_2_noclaim.json; cp gpurun_out/r2_strong_2.err gpurun_out/r2_strong_2_noclaim.err
grep -A30 "sharded results differ" gpurun_out/r2_strong_2.err | head -40
