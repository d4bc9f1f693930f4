# interleaved K1a with a CTA barrier per frame (7) / every 4 frames (8) against plain interleaving (6) and the default runs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FA_K1A_VARIANT=7 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "levels_16k or sample_rates or ragged or batch_submit" 2>&1 | tail -2
timeout 300 python profiles/stage_times.py runs= inter=FA_K1A_VARIANT:6 sync1=FA_K1A_VARIANT:7 sync4=FA_K1A_VARIANT:8 2>&1 | tail -4 | tee gpurun_out/r2o_stage_times.jsonl
for v in 7 8; do FA_K1A_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fa_fftmag -s 3 -c 1 python profiles/stage_times.py x= 2>&1 | grep -E "gpu__time|hit_rate|issue_active"; done | tee gpurun_out/r2o_ncu.txt
