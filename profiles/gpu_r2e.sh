# K8 (warp-cooperative objective): level-12 parity + stage time; host-side breakdown of the int16 / u8 e2e steps
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "level12 or levels_16k or stream_mode_every_level" 2>&1 | tail -5
LEVEL=12 WANT_SPEC=0 timeout 300 python profiles/stage_times.py l12warp= 2>&1 | tail -2
for m in u8 feat; do for d in 2 3; do timeout 200 python profiles/e2e_breakdown.py $d sink $m 2>&1 | tail -1; done; done
