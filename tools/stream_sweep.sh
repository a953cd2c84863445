#!/bin/bash
# BASELINE.json config 5: multi-stream sweep, 64-4096 concurrent 1080p camera streams on N GPUs of one box.
#   tools/stream_sweep.sh [N] > profiles/rX_stream_sweep_nN.jsonl   (N > 1: one torchrun per stream count, distinct ports)
N=${1:-1}
for S in 64 256 1024 4096; do
  if [ "$N" = "1" ]; then
    python bench.py --streams $S --steps 5 --warmup 3 --no-cpu-baseline --no-steady-state 2>/dev/null
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + S / 64)) \
      bench.py --gpus $N --streams $S --steps 5 --warmup 3 --no-cpu-baseline --no-steady-state 2>/dev/null
  fi
done
