#!/bin/bash
# C5 tuning sweep (one GPU): each line = one bench.py run of the C5 workload with the given environment.
# usage: tools/c5_sweep.sh out.log "NAME|ENV=.. ENV=.." ...
out=$1; shift
: > "$out"
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}
  line=$(env $envs timeout 600 python bench.py --workload c5 --quick --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1)
  python - "$name" "$envs" "$line" >> "$out" <<'PY'
import json, sys
name, envs, line = sys.argv[1:4]
try:
    d = json.loads(line)
    print(f"{name:28s} ms/step {d['ms_per_step']:8.3f}  sensor stage {d['stage_ms']['sensor']:8.3f}  main kernel {d['stage_ms']['sensor_main_kernel']:8.3f}  kernels {d['kernels_per_update']}  [{envs}]")
except Exception as e:
    print(f"{name:28s} FAILED {e!r} [{envs}]")
PY
done
cat "$out"
