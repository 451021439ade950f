#!/bin/bash
# Scaling runs (one box, N GPUs visible): weak scaling of the default bench at N = 1,2,4,8 and strong scaling of
# the 50k-gene synthetic (BASELINE configs[3]) at N = 1 and N = max.  Usage: profiles/scaling.sh <maxN> <tag>
MAXN=${1:-8}; TAG=${2:-r1}
OUT=gpurun_out
run() {  # N, extra args..., output name
  local N=$1; shift; local NAME=$1; shift
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + N)) bench.py --gpus $N "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
  fi
  echo "$NAME rc=$? $(tail -c 200 $OUT/$NAME.err | tr '\n' ' ')"
}
for N in 1 2 4 8; do
  [ $N -le $MAXN ] && run $N scale_weak_syn10k_n${N}_$TAG --steps 10 --warmup 3 --no-cpu-baseline
done
run 1 scale_strong_syn50k_n1_$TAG --workload syn50k --scaling strong --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
run $MAXN scale_strong_syn50k_n${MAXN}_$TAG --workload syn50k --scaling strong --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
