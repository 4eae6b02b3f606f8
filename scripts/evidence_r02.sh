mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_ref.json 2> gpurun_out/r2r_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2r_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sfx_scan_kernel -c 1 -o gpurun_out/r2r_sfx python bench.py --gib 4 --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2r_ncu_sfx.log 2>&1; echo "sfx rc=$?"
ncu --set full --clock-control none -k regex:summarize_kernel -c 1 -o gpurun_out/r2r_sum python bench.py --gib 4 --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2r_ncu_sum.log 2>&1; echo "sum rc=$?"
ncu --set full --clock-control none -k regex:kr_scan_kernel -c 1 -o gpurun_out/r2r_kr python scripts/one_scan.py 1073741824 planted kr > gpurun_out/r2r_ncu_kr.log 2>&1; echo "kr rc=$?"
ncu --set full --clock-control none -k regex:dfa_ -c 1 -o gpurun_out/r2r_dfa_small python scripts/one_scan.py 1073741824 ab auto > gpurun_out/r2r_ncu_dfa.log 2>&1; echo "dfa rc=$?"
python scripts/ref_exe_perf.py 256 > gpurun_out/r2r_exe.json 2>&1; echo "exe rc=$?"
python scripts/host_regimes.py 1024 8 > gpurun_out/r2r_regimes.log 2>&1; echo "regimes rc=$?"
scripts/microbench/host_mem 256 > gpurun_out/r2r_hostmem.json 2>&1
