#!/bin/bash
# One GPU call that settles the prepared k_bwd2 variants:
#   1. op parity suite under MT_DEBUG_FLAGS 8 / 16 / 24 (already green under the CPU emulation tier);
#   2. same-box alternating A/B of the 128-px training step (gated + MedT) for flags 0 / 8 / 16 / 24, twice;
#   3. the 512-px / L=256 configurations (throughput-bound) for the same flags.
# Usage (from the repo root, on the GPU box):  bash tools/ab_bwd_variants.sh  > gpurun_out/ab_bwd_variants.log 2>&1
set -u
cd "$(dirname "$0")/.."
echo "== parity"; MEDT_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -3
for rep in 1 2; do
  for f in 0 8 16 24; do
    echo "== bench flags=$f rep=$rep"
    MT_DEBUG_FLAGS=$f BENCH_VERBOSE=1 timeout 200 python bench.py --no-cpu-baseline --steps 30 2>&1 | grep -E "timed|Error"
  done
done
for f in 0 8 16 24; do
  echo "== 512 px flags=$f"
  MT_DEBUG_FLAGS=$f timeout 300 python tools/run_512.py 2>&1 | tail -3
done
