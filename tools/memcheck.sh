# compute-sanitizer memcheck over the small-size GPU tests of the kernels added after the forward path
mkdir -p gpurun_out
run() { echo "== $*"; timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "$@" -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error|error" | head -8 | cut -c1-220; }
run tests/test_tsdf.py -k "golden"
run tests/test_sphere_feature.py
run tests/test_sweep.py -k "rays_to_images or volume_merge"
run tests/test_backward.py -k "edge_sizes and 1-fp32"
run tests/test_lattice.py -k "fp32"
