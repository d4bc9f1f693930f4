# bench line with the per-kernel table (fa_spectrum_split_times) and the dominant-kernel roofline; C3 path smoke; new test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_times or batch_submit or levels_16k" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 300 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
timeout 400 python bench.py --workload c3 --no-cpu-baseline --steps 5 > gpurun_out/r2_c3_n1.json 2> gpurun_out/r2_c3_n1.err
tail -c 200 gpurun_out/r2_c3_n1.json; tail -3 gpurun_out/r2_c3_n1.err
