# compute-sanitizer on what changed in session 3: the sqrt fast path (2048 / generic / big FFT kernels, incl. the silence fall-back),
# fa_segment2_kernel on its slim shared-memory slice (13.5 KB per warp; 128-register build forced), the general kernel behind it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='sqrt_fast or (fft_size_sweep and (512 or 4096)) or (dense_peak and 5.0-0]) or (cuda_path and synth_sr16000_seed1_u0_l13 and not 0p and not 1-synth) or (levels_16k and 13 and not 0p and not 1])'
FA_K3_REGS=128 timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitizer3_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer3_memcheck.txt; tail -4 gpurun_out/r2_sanitizer3_memcheck.txt
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sqrt_fast or (fft_size_sweep and 0.8-4096) or (levels_16k and 13-0])" > gpurun_out/r2_sanitizer3_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer3_racecheck.txt; tail -4 gpurun_out/r2_sanitizer3_racecheck.txt
