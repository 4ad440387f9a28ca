#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over smoke() and the overlap / crowd parity tests; logs -> gpurun_out/
export EOT_SANITIZE=1
run() { # tool, label, command...
  tool=$1; label=$2; shift 2
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 "$@" > gpurun_out/sanitize_${tool}_${label}.log 2>&1
  echo "$tool $label rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${label}.log | tail -1)"
}
run memcheck smoke python -c "import __graft_entry__ as g; g.smoke()"
run memcheck tests python -m pytest -x -q tests/test_gpu_backward.py::test_backward_many_mutually_overlapping_boxes "tests/test_gpu_crowd_and_sizes.py::test_crowd_on_one_image_forward_bit_exact_backward_rel_l2[40]" tests/test_gpu_crowd_and_sizes.py::test_total_boxes_is_a_capacity tests/test_gpu_forward.py::test_forward_perspective_row
run racecheck smoke python -c "import __graft_entry__ as g; g.smoke()"
run racecheck tests python -m pytest -x -q tests/test_gpu_backward.py::test_backward_many_mutually_overlapping_boxes
run synccheck smoke python -c "import __graft_entry__ as g; g.smoke()"
