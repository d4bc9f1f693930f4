# does the segment scan at 128 registers (launch bound 480: no spills) fit beside a half-grid K1a (one CTA per SM) so that the
# scan of one batch hides behind the FFT of the next?  resident ms/step of C2, 4 batches in flight
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python profiles/stage_times.py serial= k3r136=FA_K3_REGS:136 k3r136w1=FA_K3_REGS:136,FA_K3_WARPS:1 half=FA_K1A_VARIANT:5 2>&1 | tail -4 | tee gpurun_out/r2j_stage_times.jsonl
for v in "" "FA_K3_REGS=136" "FA_K3_REGS=136 FA_K1A_VARIANT=5" "FA_K3_REGS=136 FA_K1A_VARIANT=5 FA_K3_WARPS=1" "FA_K3_REGS=136 FA_K3_WARPS=1" "FA_K3_REGS=136 FA_K1A_VARIANT=5 FA_K3_PRIO=1" "FA_K1A_VARIANT=5"; do
  env $v timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms/step', round(d['ms_per_step'],4), 'spectrum', round(d['stages']['spectrum']['ms'],3), 'segment', round(d['stages']['segment']['ms'],3))"
done | tee gpurun_out/r2j_overlap3.txt
