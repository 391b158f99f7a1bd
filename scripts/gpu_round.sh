#!/bin/bash
# One GPU-box session: parity tests, bench line, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tests|bench|ncu ...]
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log ;;
    ops)
      timeout 600 python scripts/profile_ops.py --dbg ${DBGS:-0,1,2,4} --out gpurun_out/ops.tsv 2>&1 | tail -80 | tee gpurun_out/ops.log ;;
    opsnb)
      timeout 600 python scripts/profile_ops.py --kind nb --dbg ${DBGS:-0} --out gpurun_out/ops_nb.tsv 2>&1 | tail -60 | tee gpurun_out/ops_nb.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log ;;
    bench)
      nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv | tee gpurun_out/gpu.csv
      timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
      tail -5 gpurun_out/bench.err ;;
    benchnb)
      timeout 900 python bench.py --kind nb --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_nb.err | tee gpurun_out/bench_nb.json ;;
    benchx3)
      timeout 900 python bench.py --precision fp16x3 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_x3.err | tee gpurun_out/bench_x3.json ;;
    refarm)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json ;;
    sweep)
      for cs in 1 2 4; do
        echo "== HUDIFF_CLUSTER=$cs"
        HUDIFF_CLUSTER=$cs timeout 600 python bench.py --steps 2 --warmup 3 --T 10 --no-cpu-baseline 2> gpurun_out/sweep_$cs.err \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['forward_breakdown_ms'], d['roofline']['achieved'])"
        tail -3 gpurun_out/sweep_$cs.err
      done ;;
    twocta)
      for v in 0 1 0 1; do
        echo "== HUDIFF_2CTA=$v"
        HUDIFF_2CTA=$v timeout 600 python bench.py --steps 3 --warmup 3 --T 10 --no-cpu-baseline 2> gpurun_out/twocta_$v.err \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['forward_breakdown_ms'], d['roofline']['achieved'], d['clocks'])"
        tail -3 gpurun_out/twocta_$v.err
      done ;;
    ncudbg)
      for d in ${DBGS:-1 2}; do
        HUDIFF_DBG=$d timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1224 -c 408 --csv \
          --log-file gpurun_out/launches_dbg$d.csv python bench.py --steps 1 --warmup 3 --T 4 --no-cpu-baseline \
          > gpurun_out/ncu_bench_dbg$d.log 2>&1
      done ;;
    dbg)
      for d in 0 1 2 3 4 5 7; do
        echo "== HUDIFF_DBG=$d"
        HUDIFF_DBG=$d timeout 600 python bench.py --steps 1 --warmup 3 --T 4 --no-cpu-baseline 2> gpurun_out/dbg_$d.err \
          | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['forward_breakdown_ms'])"
        tail -2 gpurun_out/dbg_$d.err
      done ;;
    ncu)
      # launch list of one timed loop (T=4 -> 4 steps of 104 launches; 3 warm-up loops = 1248 launches skipped)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1248 -c 416 --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --T 4 --no-cpu-baseline \
        > gpurun_out/ncu_bench.log 2>&1
      tail -2 gpurun_out/ncu_bench.log
      # full capture of 20 consecutive GEMM launches of the first timed step (67 GEMMs per step, 804 skipped):
      # the six wide ByteNet blocks (PFF1 / dilated conv / PFF2), then the first QKV projection and out-projection
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s ${NCU_SKIP:-822} -c ${NCU_COUNT:-20} \
        -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --T 4 --no-cpu-baseline \
        > gpurun_out/ncu_full.log 2>&1
      tail -2 gpurun_out/ncu_full.log
      # and the persistent attention kernel
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 30 -c 1 \
        -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --T 4 --no-cpu-baseline \
        > gpurun_out/ncu_attn.log 2>&1
      tail -2 gpurun_out/ncu_attn.log
      ls -la gpurun_out ;;
  esac
done
