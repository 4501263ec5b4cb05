python -m pytest tests -m gpu -q -rA -k "ray_parallel or sharded_adan" > gpurun_out/r2e_pytest.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2e_bench_n2_rays.json 2> gpurun_out/r2e_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 3 --render-split views > gpurun_out/r2e_bench_n2_views.json 2>> gpurun_out/r2e_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 12 --warmup 3 --timeline > gpurun_out/r2e_timeline_n2_rays.json 2>> gpurun_out/r2e_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 12 --warmup 3 --timeline --render-split views > gpurun_out/r2e_timeline_n2_views.json 2>> gpurun_out/r2e_n2.err
tail -4 gpurun_out/r2e_pytest.txt
