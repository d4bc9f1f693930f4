# K1b wave quantisation: kernel durations (ncu, time only) for 592 / 1000 / 1184 utterances per batch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 592 1000 1184; do
  N_UTT=$n timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fa_smooth_bands|fa_fftmag|fa_segment2|fa_peaks2" -s 12 -c 8 --csv --log-file gpurun_out/k1b_tail_$n.csv python profiles/stage_times.py x= > /dev/null 2>&1
  echo "N_UTT=$n"; python profiles/launch_summary.py gpurun_out/k1b_tail_$n.csv | head -6
done
