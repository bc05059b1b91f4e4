#!/usr/bin/env bash
# compute-sanitizer pass over the small-shape GPU tests (race / memory checks on tiny graphs).
# One tool per invocation (memcheck | racecheck | synccheck | initcheck); run it only after the same pytest selection
# has passed without the sanitizer.  Example (GPU box):
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck > gpurun_out/sanitize_memcheck.log 2>&1; tail -5 gpurun_out/sanitize_memcheck.log'
set -euo pipefail
tool="${1:-memcheck}"
sel="${2:-spmm or biagg or transr or bpr or adam or softmax or topk or sampler or compact or peer_push}"
cd "$(dirname "$0")/.."
exec compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 0 \
    python -m pytest tests -m gpu -x -q -k "$sel and not c3 and not full_size and not shape"
