#!/bin/bash
# compute-sanitizer pass over a reduced-size subset of the GPU tests (the reference relies on ASan-less gtests;
# the hand-rolled mbarrier / TMEM / TMA pipelines here need their own race and bounds checks).
#   memcheck  : out-of-bounds / misaligned global + shared accesses
#   racecheck : shared-memory hazards (the cp.async.bulk ring of the fused IK kernel, the stage rings of the tcgen05 kernels,
#               the cp.async ring / tile storage of ik_solve_mma_kernel, the operand ring of ik_poseblend_tc_kernel)
#   synccheck : divergent / invalid barrier use (named barriers per frame team)
# Output: gpurun_out/sanitize_<tool>.log and a one-line-per-tool summary gpurun_out/sanitize_summary.txt
mkdir -p gpurun_out
SEL='test_forward_vs_reference_golden or test_model_skinning_vs_oracle or test_ik_step_vs_reference_golden or test_ik_step_vposer or test_shared_beta_single or test_closest_points_vs_oracle or test_decoder_vs_reference_golden or test_body_stage_every_iteration or test_config4_vposer_every_iteration or test_shared_beta_16_frames or test_sweep_grid_bounds'
: > gpurun_out/sanitize_summary.txt
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" == "racecheck" ] && extra="--racecheck-report analysis"
  timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --error-exitcode 99 --print-limit 400 \
      python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  errs=$(grep -c "========= \(Error\|ERROR\|Invalid\|Race\|Barrier error\|Warning: Race\)" gpurun_out/sanitize_$tool.log)
  summ=$(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -1)
  tests=$(grep -E "passed|failed" gpurun_out/sanitize_$tool.log | tail -1)
  echo "$tool rc=$rc reports=$errs | $summ | $tests" | tee -a gpurun_out/sanitize_summary.txt
done
