# compute-sanitizer memcheck, broader pass: the golden parity tests (all three precision modes, both datasets), the tile-program
# layer tests, the latent-table tests on the small goldens, backward / TSDF / sphere-resampling / decoder tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { echo "== $*"; timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "$@" -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error|error|Timeout|timeout" | head -8 | cut -c1-220; echo "rc=${PIPESTATUS[0]}"; }
run tests/test_gpu_parity.py -k "not full_size and not extreme and not large"
run tests/test_gpu_tc_layers.py tests/test_gpu_fp32tc.py -k "layer_by_layer or bit_identical or bit_equal"
run tests/test_gpu_preproj.py -k "mini or identity or adversarial_bf"
run tests/test_backward.py tests/test_tsdf.py tests/test_sphere_feature.py tests/test_decoder.py -k "not large and not full"
