#!/bin/bash
# Condense an .ncu-rep (ncu --set full) into the lines the round summaries quote.  Usage: tools/ncu_summary.sh gpurun_out/x.ncu-rep > profiles/x.txt
rep="$1"
echo "# $(basename "$rep") -- ncu --set full --clock-control none (per-launch; cold-cache, serialised replays: shares, not absolutes)"
ncu -i "$rep" --page details 2>/dev/null | grep -E "^  [a-zA-Z<]|Duration|SM Frequency|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Executed Ipc Active|Issue Slots Busy|No Eligible|Registers Per|Dynamic Shared Memory Per Block|Theoretical Occupancy|Achieved Occupancy|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Waves Per SM"
echo
echo "# per launch (raw page): DRAM bytes, duration, instructions, tensor-pipe activity (sm__pipe_tensor*: share of cycles the tensor pipe was active)"
ncu -i "$rep" --page raw --csv 2>/dev/null | python3 "$(dirname "$0")/ncu_raw.py"
