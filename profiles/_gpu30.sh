#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t30.log
for b in 8 16 32 64; do for sw in "LDM_B200_LEAN_EW4=1" "LDM_B200_LEAN_EW4=0"; do env AB_B=$b $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab30.log 2>&1; done; done
tail -3 gpurun_out/r2_t30.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab30.log
