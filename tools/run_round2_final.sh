# Final round-2 evidence of the current tree on ONE B200 (run through gpurun; results land in gpurun_out/, the ones to keep
# are copied / summarised into profiles/ afterwards).   bash tools/run_round2_final.sh <tag>
T=${1:-r02F}
O=gpurun_out
set -x
python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log; tail -2 $O/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python bench.py > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err; tail -c 300 $O/${T}_bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_c2_reference_arm.json 2>&1
python bench.py --N 16 --K 16 --D 8 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
python bench.py --N 784 --K 10 --D 5 --batch 100000 --steps 10 --warmup 3 --no-cpu-baseline --no-sweep > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
python bench.py --dtype complex64 --steps 50 --warmup 5 --no-cpu-baseline --no-sweep > $O/${T}_bench_c2_c64.json 2> $O/${T}_bench_c64.err
python bench.py --N 16 --K 16 --D 8 --dtype complex64 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/${T}_bench_c3_c64.json 2> $O/${T}_bench_c3_c64.err
python tools/bench_degree.py > $O/${T}_bench_degree.json 2>&1
python tools/e2e_breakdown.py > $O/${T}_e2e_breakdown.txt 2>&1
python tools/peak_evidence.py > $O/${T}_peaks.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_bench_c2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_launch.log 2>&1
for cfg in "4 4 3 1000000 c2" "8 8 1 1000000 c5d1" "8 8 16 1000000 c5d16" "16 16 8 1000000 c3" "784 10 5 100000 c4"; do set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:qkan_block --launch-skip 2 -c 1 -o $O/${T}_$5 -f python tools/run_one.py $1 $2 $3 $4 > $O/${T}_ncu_$5.log 2>&1
done
ls -la $O/${T}_*
