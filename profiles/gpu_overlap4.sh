# half-grid K1a (FA_K1A_VARIANT=5: one CTA per SM, the other half of the register file stays free for other batches' kernels):
# resident ms/step of C2 against the number of batches in flight, sub-batches and the K3 priority stream
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { env $1 timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 30 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $2', 'ms/step', round(d['ms_per_step'],4))"; }
( run "" ""; run "FA_K1A_VARIANT=5" ""; run "" ""; run "FA_K1A_VARIANT=5" "";
  for d in 2 3 5 6 8; do run "FA_K1A_VARIANT=5" "--depth $d"; done
  for d in 3 6; do run "" "--depth $d"; done
  run "FA_K1A_VARIANT=5 FA_K3_PRIO=1" ""; run "FA_K1A_VARIANT=5" "--pipeline 2"; run "FA_K1A_VARIANT=5" "--pipeline 2 --depth 2" ) | tee gpurun_out/r2k_overlap4.txt
