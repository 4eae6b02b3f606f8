mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2zm_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2zm_pytest.log
python bench.py > gpurun_out/r2zm_bench_n1.json 2> gpurun_out/r2zm_bench_n1.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:deep_scan_kernel -c 1 -o gpurun_out/r2zm_deep python scripts/one_scan.py 536870912 almost auto > gpurun_out/r2zm_ncu_deep.log 2>&1; echo "deep rc=$?"
ncu --set full --clock-control none -k regex:dfa_ -c 1 -o gpurun_out/r2zm_dfa_small python scripts/one_scan.py 1073741824 ab auto > gpurun_out/r2zm_ncu_dfa.log 2>&1; echo "dfa rc=$?"
python scripts/host_regimes.py 1024 12 > gpurun_out/r2zm_regimes.log 2>&1; echo "regimes rc=$?"
