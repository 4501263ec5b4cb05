# 8-GPU box: the bench at N = 8 (ray-parallel) with the per-rank timeline of 8 extra steps.  One launch only: an 8-GPU box is charged 8x.
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 12 --warmup 3 --timeline-out gpurun_out/r2p_timeline_n8_rays.json > gpurun_out/r2p_bench_n8_rays.json 2> gpurun_out/r2p_n8.err
echo "rc=$?"; tail -c 400 gpurun_out/r2p_bench_n8_rays.json
