# Round-2 evidence on one B200 (run through gpurun); outputs under gpurun_out/, summarised into profiles/ by
# scripts/ncu_summary.py and scripts/ncu_kernel_md.py.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2z_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none -k regex:kr_scan_kernel -c 1 -o gpurun_out/r2z_kr python scripts/one_scan.py 1073741824 planted kr > gpurun_out/r2z_ncu_kr.log 2>&1; echo "kr rc=$?"
ncu --set full --clock-control none -k regex:dfa_ -c 1 -o gpurun_out/r2z_dfa_small python scripts/one_scan.py 1073741824 ab auto > gpurun_out/r2z_ncu_dfa.log 2>&1; echo "dfa rc=$?"
python scripts/host_regimes.py 1024 12 > gpurun_out/r2z_regimes.log 2>&1; echo "regimes rc=$?"
