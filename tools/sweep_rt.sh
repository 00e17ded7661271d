# A/B on one box: task distribution of the walk kernel (-1: fixed interleaved shares)
for rep in 1 2; do
for t in -1 0 25 50; do
  echo "WALK_TAIL_PCT=$t"
  AMCLJ_WALK_TAIL_PCT=$t python bench.py --quick --workload c3 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['ms_per_step'], d['stage_ms']['sensor'])"
done; done
for t in -1 25; do
  echo "WALK_TAIL_PCT=$t"
  AMCLJ_WALK_TAIL_PCT=$t AMCLJ_RAYTABLE=0 python bench.py --quick --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 walk', d['ms_per_step'], d['stage_ms']['sensor'])"
done
