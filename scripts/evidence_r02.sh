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
# kernel captures of the forward walkers (profiles/r02_deep_kernel.md, r02_dfa_small_kernel.md) and the host regimes
ncu --set full --clock-control none --import-source on -k regex:deep_scan_kernel -c 1 -o gpurun_out/r2z_deep python scripts/one_scan.py 536870912 almost auto > gpurun_out/r2z_ncu_deep.log 2>&1; echo "deep rc=$?"
ncu --set full --clock-control none -k regex:summarize_kernel -c 1 -o gpurun_out/r2z_sum python bench.py --gib 4 --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2z_ncu_sum.log 2>&1; echo "sum rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sfx_scan_kernel -c 1 -o gpurun_out/r2z_sfx python bench.py --gib 4 --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2z_ncu_sfx.log 2>&1; echo "sfx rc=$?"
python scripts/ref_exe_perf.py 256 > gpurun_out/r2z_exe.json 2>&1; echo "exe rc=$?"
scripts/microbench/host_mem 256 > gpurun_out/r2z_hostmem.json 2>&1
