#!/bin/bash
# Co-residency sweep (resident throughput only): hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) x K3 on a
# high-priority stream x K1a variant x batches in flight x sub-batches.
# usage (GPU box): bash profiles/sweep_overlap.sh > gpurun_out/sweep_overlap.txt
for conn in 32; do for prio in 0 1; do for v in 0 1; do for d in 2 3; do for pl in 2 4 8; do
  out=$(CUDA_DEVICE_MAX_CONNECTIONS=$conn FA_K3_PRIO=$prio FA_K1A_VARIANT=$v python bench.py --no-e2e --no-cpu-baseline --steps 12 --warmup 3 --depth $d --pipeline $pl 2>&1 | tail -1)
  echo "conn=$conn prio=$prio k1a=$v depth=$d subbatches=$pl $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step=%.3f' % (d['ms_per_step']))" 2>&1 | tail -1)"
done; done; done; done; done
