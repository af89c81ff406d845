#!/bin/bash
# round 2, GPU call G (2 GPUs): full GPU suite with the graph-replayed V-cycle (incl. the 2-GPU NCCL tests), then N=2 strong A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --durations=4 > gpurun_out/r2g_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2g_gpu_suite.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 --scaling strong --no-e2e --no-cpu-baseline > gpurun_out/r2g_bench_n2_graph.json 2> gpurun_out/r2g_bench_n2_graph.err
echo "rc=$?" >> gpurun_out/r2g_bench_n2_graph.err
FLUCA_B200_NO_GRAPH=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 6 --warmup 3 --scaling strong --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2g_bench_n2_nograph.json 2> gpurun_out/r2g_bench_n2_nograph.err
echo "rc=$?" >> gpurun_out/r2g_bench_n2_nograph.err
tail -4 gpurun_out/r2g_gpu_suite.log; for f in graph nograph; do head -c 260 gpurun_out/r2g_bench_n2_$f.json; echo; grep "ms per step\|rc=\|rror" gpurun_out/r2g_bench_n2_$f.err | tail -4; done
