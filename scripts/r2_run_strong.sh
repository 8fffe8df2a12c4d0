# region-sharded strong-scaling bench on N GPUs (BASELINE configs[3]): bash scripts/r2_run_strong.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_strong_$N.json 2> gpurun_out/r2_strong_$N.err
echo rc=$?
tail -c 3500 gpurun_out/r2_strong_$N.json; tail -5 gpurun_out/r2_strong_$N.err
