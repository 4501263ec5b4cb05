# 1-GPU box: validate the chain-kernel changes (elected MMA issue, packed bf16 split, level pairs, ping-pong E) and time the A/B arms.
timeout 600 python -m pytest tests/test_field_gpu.py -m gpu -q -rA -x > gpurun_out/r2q_pytest_field.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA --deselect tests/test_field_gpu.py > gpurun_out/r2q_pytest_rest.txt 2>&1
for impl in tcgen05 tcgen05_single_e tcgen05_split_scatter; do echo "== impl $impl"; timeout 120 python tools/prof_render.py --impl $impl --iters 8; done > gpurun_out/r2q_prof_render.txt 2>&1
for agg in 25 100 200; do echo "== agg $agg"; timeout 120 python tools/prof_render.py --agg $agg --iters 8; done >> gpurun_out/r2q_prof_render.txt 2>&1
timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -3 gpurun_out/r2q_pytest_field.txt; tail -3 gpurun_out/r2q_pytest_rest.txt; grep -E "==|k=13" gpurun_out/r2q_prof_render.txt; tail -c 300 gpurun_out/r2q_bench.json
