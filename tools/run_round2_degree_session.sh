# Evidence of round 2's third session (the degree-evaluation path) on ONE B200, run through gpurun; results land in gpurun_out/,
# the ones to keep are copied to profiles/ by hand.      bash tools/run_round2_degree_session.sh <tag>
T=${1:-r03}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
QKAN_GRAM_MINB=4 timeout 100 python -m pytest tests/test_gpu_degree.py -q -m gpu -k "gram or evaluate or normal" >> $O/${T}_pytest_gpu.log 2>&1; echo "pytest MINB=4 rc=$?" >> $O/${T}_pytest_gpu.log
QKAN_GRAM_MINB=3 timeout 100 python -m pytest tests/test_gpu_degree.py -q -m gpu -k "gram or evaluate or normal" >> $O/${T}_pytest_gpu.log 2>&1; echo "pytest MINB=3 rc=$?" >> $O/${T}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
timeout 400 python bench.py > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 120 python tools/bench_degree.py > $O/${T}_bench_degree.json 2> $O/${T}_bench_degree.err
timeout 120 python tools/bench_residuals.py > $O/${T}_bench_residuals.jsonl 2> $O/${T}_bench_residuals.err
timeout 120 python tools/profile_degree.py > $O/${T}_profile_degree.json 2> $O/${T}_profile_degree.err
timeout 100 python tools/tune_gram_waves.py > $O/${T}_gram_waves.jsonl 2> $O/${T}_gram_waves.err
# ncu last (numbers printed under ncu are never bench values): the five kernels of one evaluate_degree
timeout 200 ncu --set full --clock-control none --import-source on -k regex:qkan_cheb -c 5 -o $O/${T}_degree -f python tools/bench_degree.py --no-cpu > $O/${T}_ncu_degree.log 2>&1
# read here with: python tools/ncu_summary.py gpurun_out/${T}_degree.ncu-rep > profiles/${T}_ncu_degree_kernels.txt
