#!/bin/bash
# Round-end measurement pass on ONE B200 (run under gpurun): GPU test suite, bench lines of every BASELINE configuration,
# the ncu launch list of one EM iteration with DRAM bytes, and full ncu captures of the three hot kernels.
mkdir -p gpurun_out/final
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/final/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/final/bench5.err | tail -1 > gpurun_out/final/bench_n1.json
for c in 1 2 3 3dsc 4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 2>gpurun_out/final/bench_$c.err | tail -1 > gpurun_out/final/bench_cfg$c.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/final/launches_step.csv python tools/profile_step.py 1000000 2 > gpurun_out/final/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|gl_state_tc|gl_row_threshold|gl_post_slice_kernel' -c 8 \
  -o gpurun_out/final/hot python tools/profile_step.py 75776 1 > gpurun_out/final/hot.log 2>&1
tail -2 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/final/bench_ref.err | tail -1 > gpurun_out/final/bench_reference_arm.json; head -c 300 gpurun_out/final/bench_reference_arm.json
