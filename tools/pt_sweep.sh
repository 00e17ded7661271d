#!/bin/bash
# look-up kernel tuning sweep (one GPU): tools/pt_sweep.sh out.log workload "NAME|ENV=.. ENV=.." ...
out=$1; shift; wl=$1; shift
: > "$out"
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}
  line=$(env $envs timeout 600 python bench.py --workload $wl --quick --steps 20 --warmup 3 2>/dev/null | grep '^{' | tail -1)
  python - "$name" "$envs" "$line" >> "$out" <<'PY'
import json, sys
name, envs, line = sys.argv[1:4]
try:
    d = json.loads(line)
    print(f"{name:22s} ms/step {d['ms_per_step']:8.4f}  sensor stage {d['stage_ms']['sensor']:8.4f}  main kernel {d['stage_ms']['sensor_main_kernel']:8.4f}  kernels {d['kernels_per_update']}  [{envs}]")
except Exception as e:
    print(f"{name:22s} FAILED {e!r} [{envs}]")
PY
done
cat "$out"
