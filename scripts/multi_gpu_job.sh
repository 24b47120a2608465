#!/bin/bash
# N-GPU checks (run on a box with N >= 2 GPUs: gpurun --gpus N --timeout 1500 -- 'bash scripts/multi_gpu_job.sh N'):
# the NCCL parity tests of the landmark-sharded window (not skipped here) and the bench line at N ranks.
N=${1:-2}
cd /root/repo
timeout 600 python -m pytest tests/test_dist_gpu.py -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 --frames 40 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 3000 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
# the reference arm under the same launcher: rank 0 alone runs and prints it, the other ranks exit 0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err
echo "reference arm rc=$?"; head -c 400 gpurun_out/bench_ref_n$N.json
