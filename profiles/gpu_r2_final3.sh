# bench line regenerated against the refreshed ncu capture (issue floors, traffic), C5 sweep with the corrected kernel_path labels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 300 gpurun_out/r2_bench_final.json
timeout 600 python profiles/sweep_c5.py > gpurun_out/r2_sweep_c5.json 2> gpurun_out/r2_sweep_c5.err
tail -c 200 gpurun_out/r2_sweep_c5.json
