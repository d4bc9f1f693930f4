# K1a with the warps of a CTA on consecutive frames (FA_K1A_VARIANT=6): parity, serial stage times, resident ms/step, L1 hit rate
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FA_K1A_VARIANT=6 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "levels_16k or sample_rates or ragged or batch_submit or c2_full or golden" 2>&1 | tail -2
timeout 200 python profiles/stage_times.py runs= inter=FA_K1A_VARIANT:6 2>&1 | tail -2 | tee gpurun_out/r2n_stage_times.jsonl
FA_K1A_VARIANT=6 timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:fa_fftmag -s 3 -c 1 python profiles/stage_times.py x= 2>&1 | grep -E "gpu__time|hit_rate|issue_active|lts__t" | tee gpurun_out/r2n_ncu.txt
run() { env $1 timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4))"; }
( run "FA_K1A_VARIANT=6"; run "" ) | tee gpurun_out/r2n_resident.txt
