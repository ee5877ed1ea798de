# compute-sanitizer memcheck over small GPU tests of the kernels added in round 2 (latent table build + blend, split-operand weight
# packing, the tile kernel in table / split mode on a few tiles, upsample+concat and the tf32 convolution)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { echo "== $*"; timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest "$@" -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error|error|Timeout|timeout" | head -8 | cut -c1-220; echo "rc=${PIPESTATUS[0]}"; }
run tests/test_gpu_preproj.py -k "bf_mini and fp16"
run tests/test_gpu_fp32tc.py -k "layer_by_layer and mlp_gaussian"
run tests/test_decoder.py -k "packed_pyramid"
