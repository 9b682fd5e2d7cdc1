#!/bin/bash
# strong scaling of the 4K / 1M-triangle workload (BASELINE config 4) over 1/2/4/8 B200 of one box
W=${1:-config4}
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 60 --warmup 10 --workload $W --no-cpu-baseline 2>/dev/null | tail -1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 60 --warmup 10 --workload $W 2>/dev/null | grep '^{' | tail -1; fi
done
