# A/B of walk-kernel builds / modes: time per pass + ncu warp instructions (8.0 M steps per group on syn10k)
for mode in "" "G2V_WALK_TILE=32" $EXTRA_MODES; do
  echo "MODE=${mode:-default}"
  env $mode python profiles/r2/tune.py walk --workloads syn10k syn20k 2>&1 | grep auto | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], 'canon' if d['canonical'] else 'visit', '%.4f ms' % d['pass_ms'])"
  for skip in 44 58; do
    env $mode ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:walk_ -s $skip -c 1 python profiles/r2/tune.py walk --workloads syn10k 2>&1 | grep -E "walk_.*kernel<|inst_executed|duration|issue_active" | sed 's/(WalkGraphPtrs.*//'
  done
done
