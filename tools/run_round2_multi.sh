# Round-2 multi-GPU evidence on one box with G GPUs (gpurun --gpus G).   bash tools/run_round2_multi.sh <tag> <G>
T=${1:-r02}
G=${2:-8}
O=gpurun_out
set -x
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "fused_gather or engine_device" > $O/${T}_pytest_multi.log 2>&1; tail -3 $O/${T}_pytest_multi.log
P=29600
for n in 2 4 8; do
  [ $n -le $G ] || continue
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > $O/${T}_bench_n$n.json 2> $O/${T}_bench_n$n.err
  tail -c 200 $O/${T}_bench_n$n.json
done
P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P tools/pcie_probe.py 32 > $O/${T}_pcie_probe_n$G.txt 2>&1
P=$((P+1)); PROBE_BIND=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P tools/pcie_probe.py 32 > $O/${T}_pcie_probe_n${G}_unbound.txt 2>&1
for m in zero_copy staged copy_in; do
  P=$((P+1)); QKAN_HOST_PATH=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-gather > $O/${T}_bench_n${G}_e2e_$m.json 2> $O/${T}_bench_n${G}_e2e_$m.err
done
P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P tools/degree_multi_gpu.py > $O/${T}_degree_n$G.txt 2>&1; tail -3 $O/${T}_degree_n$G.txt
ls -la $O/${T}_*
