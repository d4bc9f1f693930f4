# K3 v2 with the slim shared-memory slice (13.5 KB per warp) and the 128-register build for launches beyond one resident wave:
# parity, C2 stage times, C3 shard (12 500 utterances, one GPU, resident) with the automatic choice and with 153 registers forced
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "reference_js or dense or capacity or golden or c2_full or ragged" 2>&1 | tail -2
timeout 200 python profiles/stage_times.py slim= slim128=FA_K3_REGS:128 2>&1 | tail -2 | tee gpurun_out/r2m_stage_times.jsonl
for v in "" "FA_K3_REGS=255"; do
  env $v timeout 400 python bench.py --workload c3 --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'ms/step', round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['stages'].items()})"
done | tee gpurun_out/r2m_c3.txt
