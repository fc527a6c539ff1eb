#!/bin/bash
# One gpurun call of round 2: GPU tests, smoke, bench lines, sanitizers, ncu evidence.  Outputs: gpurun_out/.
# usage: scripts/gpu_round2.sh [tests] [bench] [prec] [ncu] [configs] [sanitize]   (no argument = everything but prec)
mkdir -p gpurun_out
what="${*:-tests bench ncu configs sanitize}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
python -c "import os; print('cores', os.cpu_count())"
if [[ $what == *tests* ]]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q -rs --timeout 600 ) > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
  ( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
fi
if [[ $what == *bench* ]]; then
  ( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
  ( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
  echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference.json
  timeout 300 python bench.py --steps 8 --warmup 3 --no-fuse-core --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/bench_per_mouse_core.json 2>> gpurun_out/bench.err
  V1T_ATTN_BWD=three timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/bench_three_pass.json 2>> gpurun_out/bench.err
fi
if [[ $what == *prec* ]]; then
  ( time timeout 1200 python scripts/attn_prec_experiment.py ) > gpurun_out/attn_prec.json 2> gpurun_out/attn_prec.err
  echo "prec exit $?"; tail -3 gpurun_out/attn_prec.err
fi
if [[ $what == *configs* ]]; then
  for c in franke ensemble scaled; do
    ( time timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-extras ) > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
    echo "bench $c exit $?"; tail -c 400 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err
  done
fi
if [[ $what == *sanitize* ]]; then
  # racecheck: one process per test group (a report in one group must not poison the next, see profiles/r2_sanitizer_summary.txt)
  : > gpurun_out/sanitizer_summary.txt
  for grp in "fused_attention" "golden and tiny_train and bf16x3" "dropout_masks or attention_probs or emit"; do
    tag=$(echo "$grp" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)
    timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 \
      -k "$grp" > gpurun_out/sanitizer_racecheck_$tag.log 2>&1
    echo "== racecheck -k '$grp' (exit $?)" >> gpurun_out/sanitizer_summary.txt
    grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_$tag.log | head -3 >> gpurun_out/sanitizer_summary.txt
  done
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 800 \
    -k "fused_attention or (gemm and tc) or (golden and tiny_train and bf16x3) or ts_mma or scaled or emit" > gpurun_out/sanitizer_memcheck.log 2>&1
  echo "== memcheck (exit $?)" >> gpurun_out/sanitizer_summary.txt
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | head -3 >> gpurun_out/sanitizer_summary.txt
  cat gpurun_out/sanitizer_summary.txt
fi
if [[ $what == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:'attn_bwd_pair_kernel|attn_bwd2_kernel|attn_fwd2_kernel' -s 12 -c 6 \
    -f -o gpurun_out/attn python bench.py --steps 1 --warmup 3 --mice 1 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/ncu_attn.log 2>&1
  echo "ncu attn exit $?"; tail -2 gpurun_out/ncu_attn.log; ls -la gpurun_out/*.ncu-rep
fi
