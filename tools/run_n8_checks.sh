# 8-GPU box: 2-rank equivalence test, then the bench at N = 8 and 4 (ray-parallel and, at 8, the round-1 view-parallel split) + timelines.
# Every multi-rank command runs under `timeout` so a collective that never completes cannot hang the box.
timeout 300 python -m pytest tests -m gpu -q -rA -k "ray_parallel" > gpurun_out/r2f_pytest.txt 2>&1
run() { n=$1; port=$2; out=$3; shift 3; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 12 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$out 2>> gpurun_out/r2f.err; echo "$out rc=$?" >> gpurun_out/r2f.err; }
run 8 29521 r2f_bench_n8_rays.json
run 8 29522 r2f_timeline_n8_rays.json --timeline
run 8 29523 r2f_bench_n8_views.json --render-split views
run 8 29524 r2f_timeline_n8_views.json --timeline --render-split views
run 4 29525 r2f_bench_n4_rays.json
tail -3 gpurun_out/r2f_pytest.txt; grep rc= gpurun_out/r2f.err
