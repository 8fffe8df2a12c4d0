set -x
mkdir -p gpurun_out
CONTEXTSV_B200_STATS=1 timeout 900 python -m pytest tests/test_dropin_cli.py -m gpu -x -q 2>&1 | tail -30
bash scripts/r2_run_strong.sh 2
