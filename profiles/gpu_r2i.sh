# K3 utterance mode vs control scan + epoch-parallel tracking (FA_K3_MODE=1) on C2, after the tracker moved to accumulate_fm2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/stage_times.py serial= mode1=FA_K3_MODE:1 mode1_w4k=FA_K3_MODE:1,FA_K3_WORKERS:4096 2>&1 | tail -3 | tee gpurun_out/r2i_stage_times.jsonl
FA_K3_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_seg -s 12 -c 6 --csv --log-file gpurun_out/r2i_mode1_launches.csv python profiles/stage_times.py x= > /dev/null 2>&1
cat gpurun_out/r2i_mode1_launches.csv | tail -8 | cut -c1-300
