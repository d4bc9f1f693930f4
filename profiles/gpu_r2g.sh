# stream-mode tracker on accumulate_fm2: parity (stream-mode tests, reference vectors in mode 1, one-hour stream) + C4 timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or c4 or long_stream or (reference_js and 1-) or ragged or dense_peak" 2>&1 | tail -5
timeout 600 python profiles/configs_c3_c4.py --c3-utts 250 > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err; tail -c 900 gpurun_out/r2g_c4.json; tail -3 gpurun_out/r2g_c4.err
