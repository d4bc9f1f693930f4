# K3 v3 (two-warp pipeline): parity on the reference vectors + stress tests, stage times vs the serial kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "0p" 2>&1 | tail -6
timeout 300 python profiles/stage_times.py serial= pipe=FA_K3_IMPL:3 2>&1 | tail -2
