set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_segment2_kernel -s 3 -c 1 -o gpurun_out/r2b_k3v2 python profiles/stage_times.py new= > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_ncu.log
ls -la gpurun_out
