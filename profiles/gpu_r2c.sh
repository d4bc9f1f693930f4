# round 2, session 2: level 12 + K2 v2 -- full GPU parity suite, ncu full capture of K2 v2
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_gpu_tests.txt
cat gpurun_out/r2c_gpu_tests.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_peaks2_kernel -s 3 -c 1 -o gpurun_out/r2c_k2v2 python profiles/stage_times.py v2= > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
