#!/bin/bash
# ncu evidence of a round (GPU box, 1 GPU): launch list of the bench step + `--set full` captures of the top kernels.
# Reports land in gpurun_out/; condense them here with tools/ncu_summary.sh / tools/summarize_launches.py into profiles/.
R=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python bench.py --launch-list --steps 2 --warmup 3 > gpurun_out/${R}_launches.log 2>&1
# render: forward (1st launch), full 13-evaluation backward (launch 9: the SDS-style backward enqueues 9 chunk launches first, 8 of them empty)
ncu --set full --clock-control none --import-source on -k regex:k_field_fwd_tc -c 1 -o gpurun_out/${R}_field_fwd -f python tools/prof_render.py --iters 1 > gpurun_out/${R}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_field_bwd_tc -s 9 -c 1 -o gpurun_out/${R}_field_bwd -f python tools/prof_render.py --iters 1 > gpurun_out/${R}_ncu2.log 2>&1
# SD: the first tile launches are the 512x512 / 256x256 VAE convolutions, launches 60.. are U-Net projections
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 1 -c 6 -o gpurun_out/${R}_gemm_vae -f python tools/prof_sd.py 1 > gpurun_out/${R}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 60 -c 8 -o gpurun_out/${R}_gemm -f python tools/prof_sd.py 1 > gpurun_out/${R}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gn_|k_flash_attn" -s 30 -c 6 -o gpurun_out/${R}_gn_attn -f python tools/prof_sd.py 1 > gpurun_out/${R}_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
