# round 2, last code state (session 3): both bench arms, ncu launch list of the bench command, ncu --set full of every kernel of a
# serial C2 step (-> profiles/r2_ncu_kernels.json via profiles/ncu_kernels.py), the C5 sweep, the C3 shard on one GPU, smoke(), GPU tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 600 gpurun_out/r2_bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fa_ -s 24 -c 8 -o gpurun_out/r2_kernels python profiles/stage_times.py x= > gpurun_out/r2_kernels.log 2>&1
tail -3 gpurun_out/r2_kernels.log
timeout 600 python profiles/sweep_c5.py > gpurun_out/r2_sweep_c5.json 2> gpurun_out/r2_sweep_c5.err
timeout 400 python bench.py --workload c3 --no-cpu-baseline --steps 5 > gpurun_out/r2_c3_n1.json 2> gpurun_out/r2_c3_n1.err
tail -c 400 gpurun_out/r2_c3_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests.txt
cat gpurun_out/r2_gpu_tests.txt
