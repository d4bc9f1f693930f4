# K1a with a chunk queue (FA_K1A_VARIANT=6 / 7: 8 / 16 frames per grab) instead of one static run per warp: serial stage time,
# parity, resident ms/step with 4 batches in flight
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FA_K1A_VARIANT=6 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "levels_16k or sample_rates or ragged or batch_submit or c2_full" 2>&1 | tail -2
timeout 200 python profiles/stage_times.py static= dyn8=FA_K1A_VARIANT:6 dyn16=FA_K1A_VARIANT:7 2>&1 | tail -3 | tee gpurun_out/r2l_stage_times.jsonl
run() { env $1 timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 30 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $2', 'ms/step', round(d['ms_per_step'],4))"; }
( run "FA_K1A_VARIANT=6" ""; run "FA_K1A_VARIANT=7" ""; run "" "" ) | tee gpurun_out/r2l_overlap5.txt
