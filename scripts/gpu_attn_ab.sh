#!/bin/bash
# A/B of the attention-backward organisations on the GPU box.  Outputs: gpurun_out/ab_*.json, gpurun_out/pytest_attn.log
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "fused_attention" ) > gpurun_out/pytest_attn.log 2>&1
echo "pytest attn exit $?"; tail -15 gpurun_out/pytest_attn.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras"
for g in 1 4 8 16 32 64; do
  V1T_ATTN_BWD_GROUP=$g timeout 300 $B > gpurun_out/ab_group$g.json 2> gpurun_out/ab_group$g.err; echo "group $g exit $?"
done
V1T_ATTN_BWD=pair timeout 300 $B > gpurun_out/ab_pair.json 2> gpurun_out/ab_pair.err; echo "pair exit $?"; tail -2 gpurun_out/ab_pair.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["phases"]
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms; attn_bwd phase", round(p["attn_bwd"]["ms_per_step"], 2),
              "kernel us", round(1e3 * p["attn_bwd_kernel"]["ms_per_step"] / max(p["attn_bwd_kernel"]["scopes_per_step"], 1), 1))
    except Exception as e:
        print(f, "ERR", e)
PY
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:'attn_bwd2_kernel' -s 4 -c 2 \
    -f -o gpurun_out/attn_bwd python bench.py --steps 1 --warmup 3 --mice 1 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/ncu_attn_bwd.log 2>&1
  echo "ncu bwd exit $?"; tail -2 gpurun_out/ncu_attn_bwd.log
  V1T_ATTN_BWD=pair timeout 600 ncu --set full --import-source on --clock-control none -k regex:'attn_bwd_pair_kernel' -s 4 -c 2 \
    -f -o gpurun_out/attn_pair python bench.py --steps 1 --warmup 3 --mice 1 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/ncu_attn_pair.log 2>&1
  echo "ncu pair exit $?"; tail -2 gpurun_out/ncu_attn_pair.log
fi
( timeout 600 python -m pytest tests/test_live_reference.py -m gpu -q -s --timeout 300 ) > gpurun_out/pytest_live.log 2>&1
echo "pytest live exit $?"; grep -E "live-ref|passed|failed" gpurun_out/pytest_live.log
