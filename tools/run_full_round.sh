set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r01d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01d_pytest_gpu.log; tail -3 gpurun_out/r01d_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01d_smoke.log 2>&1; tail -1 gpurun_out/r01d_smoke.log
python bench.py > gpurun_out/r01d_bench_c2.json 2> gpurun_out/r01d_bench_c2.err; tail -c 600 gpurun_out/r01d_bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d_bench_c2_ref.json 2>&1
python bench.py --N 16 --K 16 --D 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_c3.json 2> gpurun_out/r01d_bench_c3.err
python bench.py --N 784 --K 10 --D 5 --batch 100000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_c4.json 2> gpurun_out/r01d_bench_c4.err
python bench.py --dtype complex64 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r01d_bench_c64.json 2> gpurun_out/r01d_bench_c64.err
python bench.py --N 16 --K 16 --D 8 --dtype complex64 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_c3_c64.json 2> gpurun_out/r01d_bench_c3_c64.err
for d in 1 2 4 8 16; do python bench.py --N 8 --K 8 --D $d --batch 10000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01d_bench_c5_d${d}_n1_10M.json 2> gpurun_out/r01d_bench_c5_d${d}.err; done
python tools/e2e_ab.py > gpurun_out/r01d_e2e_ab.txt 2>&1; python tools/e2e_ab.py 784 10 5 100000 >> gpurun_out/r01d_e2e_ab.txt 2>&1
python tools/tune.py --configs c2,c5d1,c5d2,c5d4,c5d8,c5d16,c3,c4 --variants 0:0:0:0 > gpurun_out/r01d_tune_final.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_bench_c2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkan_block_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_c2 -f python tools/run_one.py 4 4 3 1000000 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkan_block_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_c3 -f python tools/run_one.py 16 16 8 50000 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkan_block_window -s 2 -c 1 -o gpurun_out/r01d_prof_c4 -f python tools/run_one.py 784 10 5 20000 > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkan_block_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_c5d16 -f python tools/run_one.py 8 8 16 100000 > gpurun_out/ncu_c5d16.log 2>&1
ls -la gpurun_out/r01d_*
