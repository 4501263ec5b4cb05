# 1-GPU box, final validation of a round: full GPU test suite, smoke, bench (with the CPU baseline), in-pipeline kernel breakdown, ncu evidence.
timeout 1200 python -m pytest tests -m gpu -q -rA > gpurun_out/r2s_pytest_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 300 python bench.py --defer-backward --no-cpu-baseline > gpurun_out/r2s_bench_deferred.json 2>> gpurun_out/r2s_bench.err
timeout 300 python bench.py --timeline --steps 8 > gpurun_out/r2s_timeline_n1.json 2>> gpurun_out/r2s_bench.err
timeout 300 python tools/kernel_breakdown.py > gpurun_out/r2s_kernel_breakdown.txt 2>> gpurun_out/r2s_bench.err
timeout 300 python tests/measure_ref_kernels.py > gpurun_out/r2s_ref_kernels.txt 2>> gpurun_out/r2s_bench.err
timeout 900 bash tools/run_profiles.sh r2s > gpurun_out/r2s_profiles.log 2>&1
grep -E "passed|failed" gpurun_out/r2s_pytest_gpu.txt | tail -1; tail -1 gpurun_out/r2s_smoke.txt; tail -c 250 gpurun_out/r2s_bench.json; head -3 gpurun_out/r2s_kernel_breakdown.txt
