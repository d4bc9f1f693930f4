set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests.txt
cat gpurun_out/r2a_gpu_tests.txt
timeout 300 python profiles/stage_times.py new= old=FA_K3_IMPL:1 new96=FA_K3_REGS:96 new4w=FA_K3_WARPS:4 > gpurun_out/r2a_stage_times.jsonl 2>gpurun_out/r2a_stage_times.err
cat gpurun_out/r2a_stage_times.jsonl; tail -5 gpurun_out/r2a_stage_times.err
which node nodejs deno bun qjs 2>&1 | head; nproc; free -g | head -2; lscpu | head -20 > gpurun_out/r2a_lscpu.txt
