# does a fine-grained (non-persistent) K1a let the segment scan of other batches run beside it?  resident ms/step, 4 batches in flight
cd $GRAFT_REPO_ROOT
for v in "" "FA_K1A_RPW=8" "FA_K1A_RPW=16" "FA_K1A_RPW=32" "FA_K1A_RPW=16 FA_K3_PRIO=1" "FA_K1A_RPW=32 FA_K3_PRIO=1" "FA_K3_PRIO=1"; do
  env $v timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms/step', round(d['ms_per_step'],4), 'spectrum', round(d['stages']['spectrum']['ms'],3), 'segment', round(d['stages']['segment']['ms'],3))"
done
