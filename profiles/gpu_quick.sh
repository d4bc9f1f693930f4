# quick GPU iteration: K3-related parity tests + per-stage times of the C2 batch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${TESTS:-reference_js or dense_peak or levels_16k or c2_full or ragged}" 2>&1 | tail -8
timeout 300 python profiles/stage_times.py ${VARIANTS:-new=} 2>&1 | tail -6
