cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c4 and not reference_js and not dense" 2>&1 | tail -4
FA_K1_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "levels_16k or sample_rates or fft_size_sweep or config_variants or ragged or golden or batch_submit" 2>&1 | tail -3
timeout 300 python profiles/stage_times.py spec_auto= spec_fused=FA_K1_FUSED:1 2>&1 | tail -3
WANT_SPEC=0 timeout 300 python profiles/stage_times.py nospec_auto= nospec_two=FA_K1_FUSED:0 2>&1 | tail -3
