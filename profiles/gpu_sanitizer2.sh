# compute-sanitizer on the kernels added after the first sanitizer pass: K8 (level 12), level-3 export, K4 staging, truncate switch,
# the stream-mode tracker on accumulate_fm2, K3 v3 (memcheck only: its shared-memory ring is polled on purpose)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='level12 or level3 or live_session or truncate or (levels_16k and not 0p) or (stream_mode_every_level and (12 or 3 or 13)) or (config_variants and (kw10 or kw13))'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL or (reference_js and 0p and synth_sr16000_seed1)" > gpurun_out/r2_sanitizer2_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer2_memcheck.txt; tail -4 gpurun_out/r2_sanitizer2_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer2_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer2_racecheck.txt; tail -4 gpurun_out/r2_sanitizer2_racecheck.txt
timeout 400 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "level12 or level3 or truncate" > gpurun_out/r2_sanitizer2_initcheck.txt 2>&1
echo "initcheck rc=$?" >> gpurun_out/r2_sanitizer2_initcheck.txt; tail -4 gpurun_out/r2_sanitizer2_initcheck.txt
