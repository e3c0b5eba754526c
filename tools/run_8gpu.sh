set -x
P=29600
for d in 1 8 16; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --N 8 --K 8 --D $d --batch 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r01d_bench_c5_d${d}_n8.json 2> gpurun_out/r01d_bench_c5_d${d}_n8.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r01d_bench_c2_n8.json 2> gpurun_out/r01d_bench_c2_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/degree_multi_gpu.py > gpurun_out/r01d_degree_n8.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/degree_multi_gpu.py > gpurun_out/r01d_degree_n2.txt 2>&1
tail -2 gpurun_out/r01d_degree_n8.txt gpurun_out/r01d_degree_n2.txt
