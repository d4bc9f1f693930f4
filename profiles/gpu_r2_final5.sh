# last code state (K1a: one 16-warp CTA per SM, window table in shared memory): GPU tests, bench line, ncu launch list, ncu --set full
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests.txt
tail -2 gpurun_out/r2_gpu_tests.txt
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 200 gpurun_out/r2_bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_ -s 24 -c 8 -o gpurun_out/r2_kernels python profiles/stage_times.py x= > gpurun_out/r2_kernels.log 2>&1
tail -2 gpurun_out/r2_kernels.log
