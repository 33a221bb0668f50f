#!/bin/bash
# Developer tool (GPU box): run the small parity cases under compute-sanitizer.  memcheck catches out-of-bounds
# reads of the mirrors (attribute table, visited bitmaps, ELL rows); racecheck the shared-memory queue merges of
# the graph kernels and the candidate buffers of bf_select_kernel; synccheck the mbarrier / __syncthreads use.
# Slow (10-50x): keep to the golden-sized tests.  Usage: gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'
set -u
tool="${1:-memcheck}"
sel="${2:-dense_vector or halfcircle or rand2k or empty_and_edge or pair_distances or normalize}"
mkdir -p gpurun_out
timeout 850 compute-sanitizer --tool "$tool" --error-exitcode 9 --log-file "gpurun_out/sanitize_${tool}.log" \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$sel" 2>&1 | tail -3
echo "sanitizer exit: $?"
tail -5 "gpurun_out/sanitize_${tool}.log"
