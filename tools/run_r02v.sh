T=${1:-r02v}
O=gpurun_out
python tools/e2e_ab.py > $O/${T}_e2e_ab.txt 2>&1; python tools/e2e_ab.py 784 10 5 100000 >> $O/${T}_e2e_ab.txt 2>&1; python tools/e2e_ab.py 16 16 8 1000000 >> $O/${T}_e2e_ab.txt 2>&1
python tools/ab_libs.py --configs c4,c3s --libs prev=gpurun_tmp/libqkan_sw16.so,new=qkan_implementation_b200/libqkan_b200.so > $O/${T}_ab.jsonl 2>&1
QKAN_ELEM_RUNTIME_G=1 python tools/ab_libs.py --configs c4 --libs new_runtimeG=qkan_implementation_b200/libqkan_b200.so >> $O/${T}_ab.jsonl 2>&1
timeout 1000 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log; tail -2 $O/${T}_pytest_gpu.log
python bench.py --no-cpu-baseline --no-sweep > $O/${T}_bench_c2_nosweep.json 2> $O/${T}_bench.err; tail -c 1500 $O/${T}_bench_c2_nosweep.json
cat $O/${T}_e2e_ab.txt; cut -c1-140 $O/${T}_ab.jsonl
