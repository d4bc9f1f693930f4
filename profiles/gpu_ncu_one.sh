# ncu --set full capture of one kernel: KERNEL=regex OUT=name [SKIP=n] [ENVV="A=1 B=2"]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
env $ENVV timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-3} -c 1 -o gpurun_out/$OUT python profiles/stage_times.py x= > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
