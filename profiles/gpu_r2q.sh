# K1a with the window table in shared memory: two 8-warp CTAs per SM (7) / one 16-warp CTA per SM sharing all tables (8)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 7 8; do FA_K1A_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "levels_16k and 5-0]" 2>&1 | tail -1; done
timeout 300 python profiles/stage_times.py base= win=FA_K1A_VARIANT:7 win16=FA_K1A_VARIANT:8 base2= 2>&1 | tail -4 | tee gpurun_out/r2q_stage_times.jsonl
for v in 7 8; do FA_K1A_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fa_fftmag -s 3 -c 1 python profiles/stage_times.py x= 2>&1 | grep -E "gpu__time|hit_rate|issue_active"; done | tee gpurun_out/r2q_ncu.txt
