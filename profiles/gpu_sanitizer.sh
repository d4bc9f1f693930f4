# compute-sanitizer over the GPU parity suite (subset sized to fit the time box); logs -> gpurun_out/r2_sanitizer_*.txt
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='levels_16k or ragged or dropped_segment or dense_peak or error_behaviour or submit_frames or level11 or int16 or (fft_size_sweep and 0.8) or (stream_mode_every_level) or (reference_js and (synth_sr16000_seed1 or voiced_random_seed1_ or adversarial_seed1_))'
for tool in memcheck racecheck initcheck synccheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.txt
  tail -4 gpurun_out/r2_sanitizer_$tool.txt
done
# K2 A/B after batching the trim re-reads
timeout 300 python profiles/stage_times.py v2c= direct=FA_K2_IMPL:-1 > gpurun_out/r2c_stage_times2.jsonl 2> gpurun_out/r2c_stage_times2.err
cat gpurun_out/r2c_stage_times2.jsonl
