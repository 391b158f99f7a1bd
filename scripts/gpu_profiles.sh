#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box from the repo root); only SMALL summaries stay in gpurun_out/
# (gpurun copies back at most 64 MiB):
#   1. launch list (device time of every kernel of a short bench run)          -> gpurun_out/r02b_launches.csv
#   2. `--set full` of the 67 GEMM launches of one sampling step                -> gpurun_out/r02b_ncu_full_gemm.json
#   3. `--set full` of one attention launch (fp16) and one of the split policy  -> r02b_attn*.ncu-rep + .json
# A number printed by a run under ncu is never a bench value.
set -u
mkdir -p gpurun_out /tmp/ncu
BENCH="python bench.py --T 4 --steps 1 --warmup 3 --no-cpu-baseline --no-x3"
if [ "${SKIP_LIST:-0}" != "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
  tail -1 gpurun_out/ncu_launches.log | cut -c1-200
fi
# one step = 67 GEMM launches; skip the first 4 steps (first warm-up loop)
timeout 1200 ncu --set full --clock-control none -k gemm_tc_kernel -s 268 -c 67 -o /tmp/ncu/r02b_gemm -f $BENCH > gpurun_out/ncu_gemm.log 2>&1
python scripts/ncu_summary.py /tmp/ncu/r02b_gemm.ncu-rep gpurun_out/r02b_ncu_full_gemm.json
timeout 600 ncu --set full --clock-control none --import-source on -k attention_tc3_kernel -s 40 -c 1 -o gpurun_out/r02b_attn -f $BENCH > gpurun_out/ncu_attn.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02b_attn.ncu-rep gpurun_out/r02b_ncu_full_attention.json
if [ "${WITH_X3:-0}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k attention_x3_kernel -s 40 -c 1 -o gpurun_out/r02b_attn_x3 -f $BENCH --precision fp16x3 > gpurun_out/ncu_attn_x3.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02b_attn_x3.ncu-rep gpurun_out/r02b_ncu_full_attention_x3.json
fi
du -sh gpurun_out
