#!/bin/bash
# compute-sanitizer runs of the tiny-sensor GPU tests (memcheck, racecheck, initcheck); logs under gpurun_out/ (copied to profiles/)
set -u
TESTS="tests/test_gpu_parity.py::test_tiny_sensor_two_sequences tests/test_gpu_parity.py::test_empty_and_ragged_inputs tests/test_gpu_parity.py::test_last_writer_wins tests/test_gpu_kf500.py::test_keyframe_map_tiny_sensor_libm tests/test_gpu_kf500.py::test_knn_ties_on_a_lattice tests/test_gpu_keyframes.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 50 --log-file gpurun_out/r2_sanitizer_$tool.log \
    python -m pytest $TESTS -m gpu -x -q > gpurun_out/r2_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_summary.txt
  tail -3 gpurun_out/r2_sanitizer_${tool}_pytest.log >> gpurun_out/r2_sanitizer_summary.txt
  grep -c "ERROR SUMMARY" gpurun_out/r2_sanitizer_$tool.log >> gpurun_out/r2_sanitizer_summary.txt
  grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log | sort | uniq -c >> gpurun_out/r2_sanitizer_summary.txt
done
