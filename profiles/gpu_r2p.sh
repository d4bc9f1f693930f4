# C3 shard (48 kHz: hop 1200, window overlap 41 %): K1a with the CTA's warps on consecutive frames (default) vs one run per warp (6)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" "FA_K1A_VARIANT=6"; do
  env $v timeout 400 python bench.py --workload c3 --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms/step', round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['stages'].items()})"
done | tee gpurun_out/r2p_c3.txt
