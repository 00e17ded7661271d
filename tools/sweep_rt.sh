# A/B on one box: ring set-up records of the walk kernel on / off, C3 and C2 (walk kernel forced)
for rep in 1 2; do
for rs in 1 0; do
  echo "RING_SETUP=$rs"
  AMCLJ_RING_SETUP=$rs python bench.py --quick --workload c3 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['ms_per_step'], d['stage_ms']['sensor'])"
  AMCLJ_RING_SETUP=$rs AMCLJ_RAYTABLE=0 python bench.py --quick --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 walk', d['ms_per_step'], d['stage_ms']['sensor'])"
done; done
