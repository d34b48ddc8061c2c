#!/bin/bash
# compute-sanitizer pass over the hand-written kernels ("Race detection / sanitizers").
# Runs on the GPU box through gpurun; small shapes only (the sanitizer slows kernels 10-100x).
#
#   memcheck   out-of-bounds / misaligned global + shared accesses (reflected row loaders, strip halos, ragged tiles)
#   racecheck  shared-memory hazards: the cp.async rings that are read back without a barrier (each lane reads only
#              its own slots), the LLG convert-once stage and G_H tile, the block reductions
#   synccheck  divergent / invalid barrier use (the tile kernels' early-outs around __syncthreads)
#   initcheck  (optional, TOOLS="... initcheck") reads of uninitialised global memory; noisy under the torch caching allocator
#
# One tool per gpurun call: TOOLS=memcheck scripts/sanitize.sh r2 (then racecheck, synccheck in their own calls).
# usage: scripts/sanitize.sh <tag>      -> gpurun_out/<tag>_sanitize_<tool>.log + gpurun_out/<tag>_sanitize_summary.txt
tag=${1:-r2}
O=gpurun_out
mkdir -p $O
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
# kernel tests on small grids, both kernel paths (fast / generic), + the slab sampler in lock step + the update kernels
SEL='heat_guidance_matches_closed_form or llg_residual_guidance or llg_norm_guidance or empty_mask or row_slab_decomposition or update_kernels_bit_exact or halo_push_and_flag_wait or update_rows or lockstep_slab or laplacian_golden or heat_loss2_golden or training'
FILES="tests/test_gpu_kernels.py tests/test_gpu_slab.py tests/test_training_loss.py"
summary=$O/${tag}_sanitize_summary.txt
echo "# compute-sanitizer over: pytest -m gpu -k \"$SEL\" $FILES" > $summary
echo "# $($SAN --version | tail -1)" >> $summary
rc_all=0
for tool in ${TOOLS:-memcheck}; do
    log=$O/${tag}_sanitize_${tool}.log
    extra=""
    [ "$tool" = "memcheck" ] && extra="--leak-check no"
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
    timeout ${SAN_TIMEOUT:-1500} $SAN --tool $tool $extra --error-exitcode 86 --print-limit 20 \
        python -m pytest $FILES -m gpu -x -q -k "$SEL" -p no:cacheprovider > $log 2>&1
    rc=$?
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
    tests=$(grep -E "passed|failed" $log | tail -1)
    echo "$tool: rc=$rc | ${errs:-no summary line} | pytest: ${tests:-no pytest line}" >> $summary
    [ $rc -ne 0 ] && rc_all=1
    # keep the logs small: the head (tool banner), every hazard / error record, the tail
    if [ $(wc -c < $log) -gt 400000 ]; then
        { head -50 $log; echo "... [trimmed] ..."; grep -E -A12 "=========( Error| Warning| Invalid| Race| Uninitialized| Barrier)" $log | head -400; echo "... [trimmed] ..."; tail -60 $log; } > $log.trim
        mv $log.trim $log
    fi
done
cat $summary
exit $rc_all
