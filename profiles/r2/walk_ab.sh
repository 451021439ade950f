# A/B of walk-kernel builds: G2VEC_B200_LIB selects the library; time per pass + ncu warp instructions (8.0 M steps per group)
for lib in "" $EXTRA_LIBS; do
  echo "LIB=${lib:-current}"
  G2VEC_B200_LIB=$lib python profiles/r2/tune.py walk --workloads syn10k syn20k 2>&1 | grep auto | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], 'canon' if d['canonical'] else 'visit', '%.4f ms' % d['pass_ms'])"
  for skip in 44 58; do
    G2VEC_B200_LIB=$lib ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:walk_kernel -s $skip -c 1 python profiles/r2/tune.py walk --workloads syn10k 2>&1 | grep -E "walk_kernel<|inst_executed|duration" | sed 's/(WalkGraphPtrs.*//'
  done
done
