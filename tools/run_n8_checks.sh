# 8-GPU box: the 2-rank equivalence tests, then the bench at N = 8 (ray-parallel) with the per-rank timeline of 8 extra steps.  An 8-GPU box is charged 8x.
timeout 300 python -m pytest tests -m gpu -q -rA -k "ray_parallel or sharded_adan" > gpurun_out/r2t_pytest_n2.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --timeline-out gpurun_out/r2t_timeline_n8_rays.json > gpurun_out/r2t_bench_n8_rays.json 2> gpurun_out/r2t_n8.err
echo "rc=$?"; tail -3 gpurun_out/r2t_pytest_n2.txt; tail -c 300 gpurun_out/r2t_bench_n8_rays.json
