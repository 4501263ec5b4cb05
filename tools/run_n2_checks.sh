# 2-GPU box: the G-rank == sequential equivalence tests, then the bench at N = 2 (ray-parallel) with the per-rank timeline of 8 extra steps.
# Every multi-rank command runs under `timeout` so a collective that never completes cannot hang the box.
timeout 400 python -m pytest tests -m gpu -q -rA -k "ray_parallel or sharded_adan" > gpurun_out/r2p_pytest_n2.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --timeline-out gpurun_out/r2p_timeline_n2_rays.json > gpurun_out/r2p_bench_n2_rays.json 2> gpurun_out/r2p_n2.err
tail -4 gpurun_out/r2p_pytest_n2.txt; tail -c 400 gpurun_out/r2p_bench_n2_rays.json
