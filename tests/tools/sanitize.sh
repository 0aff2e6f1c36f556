#!/bin/bash
# compute-sanitizer passes over the small-shape parity tests (race detection / sanitizers). Run on a GPU box:
#     gpurun --timeout 1500 -- 'bash tests/tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck (out-of-bounds / misaligned accesses), racecheck (shared-memory hazards), synccheck (barrier misuse) on the kernel-level tests
# whose shapes are small enough for a 10-100x slowdown; the TMA / tcgen05 kernels are included (memcheck follows bulk copies, racecheck
# does not see tensor-memory traffic). NOT run in rounds 1-2 (the GPU budget went into measurement): status "written, not executed".
set -u
SEL='mlp_kernel_vs_oracle or aggregate_rows_vs_oracle or blockdiag or apsp_vs_oracle_random or small_groups or local_edges or cross_entropy or adam'
for tool in memcheck racecheck synccheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 99 --launch-timeout 0 \
        python -m pytest tests -q -x -m gpu -k "$SEL" -p no:cacheprovider 2>&1 | tail -25
    echo "exit code: $?"
done
