cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_lj.py tests/test_gpu_golden.py tests/test_gpu_tracker.py tests/test_gpu_moves.py -x -q 2>&1 | tail -5 > gpurun_out/ab_tests.log
B="timeout 300 python bench.py --no-cpu-baseline --no-extra --chain 200 --steps 30 --warmup 5"
rm -f gpurun_out/ab_*.json
$B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
MCPYB_EXPAND=0 $B > gpurun_out/ab_noexpand.json 2>/dev/null
MCPYB_RB_SLICES=1 $B > gpurun_out/ab_s1.json 2>/dev/null
MCPYB_RB_SLICES=3 $B > gpurun_out/ab_s3.json 2>/dev/null
MCPYB_ROWS=pipe MCPYB_RB_SLICES=1 $B > gpurun_out/ab_pipe_s1.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lj_trial_rows_block -s 1 -c 1 -o gpurun_out/prof_r1_rows13 -f python tests/prof_dropin.py lj > gpurun_out/ncu_pipe.log 2>&1
