# synccheck / racecheck over the kernels added in round 2 (ik_solve_mma_kernel, ik_poseblend_tc_kernel, sweep grid)
mkdir -p gpurun_out
SEL='test_ik_step_vs_reference_golden or test_ik_step_vposer or test_shared_beta_16_frames or test_sweep_grid_bounds or test_task_rest_shape_kernels or random_task_sets'
for tool in synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 50 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize2_$tool.log 2>&1
  echo "$tool rc=$? | $(grep 'ERROR SUMMARY' gpurun_out/sanitize2_$tool.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitize2_$tool.log | tail -1)"
done
