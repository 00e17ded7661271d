# A/B of ray-table knobs on one box: each configuration three times, interleaved
for rep in 1 2; do
for cfg in "20" "35" "50" "100"; do
  echo "TAIL=$cfg"
  AMCLJ_RT_TAIL_PCT=$cfg python bench.py --quick --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms']['sensor'])"
done; done
