# round 2, session 2: level 3 + K2 without the old kernel / asm workarounds: full GPU parity suite, initcheck subset
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_gpu_tests.txt
cat gpurun_out/r2d_gpu_tests.txt
SEL='levels_16k or ragged or dropped_segment or error_behaviour or submit_frames or level11 or (stream_mode_every_level and 13) or (reference_js and synth_sr16000_seed1)'
timeout 500 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_initcheck.txt 2>&1
echo "initcheck rc=$?" >> gpurun_out/r2_sanitizer_initcheck.txt
grep -c "Uninitialized" gpurun_out/r2_sanitizer_initcheck.txt; tail -4 gpurun_out/r2_sanitizer_initcheck.txt
