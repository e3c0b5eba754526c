# Final round-2 multi-GPU evidence on one box with 8 GPUs (gpurun --gpus 8).   bash tools/run_round2_multi_final.sh <tag>
T=${1:-r02J}
O=gpurun_out
set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "fused_gather or engine_device" > $O/${T}_pytest_multi.log 2>&1; tail -3 $O/${T}_pytest_multi.log
P=29700
for n in 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.err
  tail -c 200 $O/${T}_bench_n$n.json
done
P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P tools/degree_multi_gpu.py > $O/${T}_degree_n8.txt 2>&1; tail -3 $O/${T}_degree_n8.txt
ls -la $O/${T}_*
