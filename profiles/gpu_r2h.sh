# round 2, session 3: sqrt fast path in every K1a kernel + single division in fm_score -- parity subset, stage times, C5 points,
# ncu of the big-FFT kernel at 4096
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sqrt_fast or fft_size_sweep or config_variants or reference_js or golden or c2_full" 2>&1 | tail -4
timeout 300 python profiles/stage_times.py r2h= 2>&1 | tail -2 | tee gpurun_out/r2h_stage_times.jsonl
for n in 4096 8192; do timeout 300 python profiles/c5_one.py $n 2>&1 | tail -1; done | tee gpurun_out/r2h_c5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fftmag_big -s 2 -c 1 -o gpurun_out/r2h_big4096 python profiles/c5_one.py 4096 > gpurun_out/r2h_big4096.log 2>&1
tail -2 gpurun_out/r2h_big4096.log
