#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench line, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 600 python -m pytest tests -m gpu -q --timeout 300 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
( time timeout 200 python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
( time timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" = "ab" ]; then  # A/B of the 128-bit readout kernels
  V1T_READOUT_V4=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_readout_scalar.json 2>> gpurun_out/bench.err
  python - <<'PY'
import json
for f in ("bench.json", "bench_readout_scalar.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, round(d["value"], 1), "samples/s", {k: round(v["ms_per_step"], 3) for k, v in d["phases"].items() if k.startswith("readout")})
PY
fi
if [ "$1" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; wc -l gpurun_out/launches.csv
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:'adamw_l1_kernel|rollout_step_kernel' -c 6 \
    -f -o gpurun_out/extras python scripts/prof_extras.py > gpurun_out/ncu_extras.log 2>&1
  echo "ncu extras exit $?"; tail -2 gpurun_out/ncu_extras.log
fi
