# one GPU call: A/B of builds, GPU tests, ncu of the C2 / C5 D1 kernels.   bash tools/run_r02t.sh <tag>
T=${1:-r02t}
O=gpurun_out
python tools/ab_libs.py --configs c2,c5d1,c5d2,c5d4,c5d8,c5d16,c3s,c4 --libs prev=gpurun_tmp/libqkan_sw16.so,new=qkan_implementation_b200/libqkan_b200.so > $O/${T}_ab.jsonl 2>&1
QKAN_DIRECT_RUNTIME_G=1 python tools/ab_libs.py --configs c2,c5d1,c5d4 --libs new_runtimeG=qkan_implementation_b200/libqkan_b200.so >> $O/${T}_ab.jsonl 2>&1
python tools/ab_libs.py --dtype complex64 --configs c2,c5d4,c3s --libs prev=gpurun_tmp/libqkan_sw16.so,new=qkan_implementation_b200/libqkan_b200.so > $O/${T}_ab_c64.jsonl 2>&1
timeout 1000 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log; tail -2 $O/${T}_pytest_gpu.log
for cfg in "4 4 3 1000000 c2" "8 8 1 1000000 c5d1"; do set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:qkan_block --launch-skip 2 -c 1 -o $O/${T}_$5 -f python tools/run_one.py $1 $2 $3 $4 > $O/${T}_ncu_$5.log 2>&1
done
cut -c1-140 $O/${T}_ab.jsonl
