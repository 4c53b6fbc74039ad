#!/bin/bash
# compute-sanitizer passes over the smoke shape (one render_image train step through every fused kernel + backward) and
# over the kernel-level GPU tests of the tensor-core kernels.  Usage (on the GPU box): bash profiles/tools/sanitize.sh OUTDIR
# Each tool's full log goes to OUTDIR/sanitizer_<tool>.log; the last lines hold the error summary.
out=${1:-gpurun_out}
mkdir -p "$out"
export PYTORCH_NO_CUDA_MEMORY_CACHING=1   # one cudaMalloc per tensor: out-of-bounds accesses cannot land in a pooled block
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ "$tool" = initcheck ] && extra="--track-unused-memory no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 40 --launch-timeout 0 \
    python -c "import __graft_entry__ as g; g.smoke()" > "$out/sanitizer_$tool.log" 2>&1
  echo "exit code $?" >> "$out/sanitizer_$tool.log"
  tail -4 "$out/sanitizer_$tool.log"
done
