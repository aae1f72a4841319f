#!/bin/bash
# compute-sanitizer passes over the parity tests of the hand-written kernels.
#   tools/sanitize.sh [outdir]      (run on the GPU box; summaries land in outdir, default gpurun_out/)
# memcheck over conv (tcgen05 + head + direct), metrics, bicubic / resize and tiling tests; racecheck and synccheck over the
# kernels that synchronise through shared memory with plain barriers (metrics, bicubic, tiling, CUDA-core conv engines).
# The tcgen05 kernels hand data between roles through mbarriers / TMA / TMEM, which racecheck does not model; they are
# covered by memcheck and by the bit-tight parity tests themselves.
out=${1:-gpurun_out}
mkdir -p "$out"
cs=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool timeout pytest-args...
  local name=$1 tool=$2 limit=$3; shift 3
  timeout "$limit" $cs --tool "$tool" --error-exitcode 99 --print-limit 20 python -m pytest -q -x -p no:cacheprovider "$@" > "$out/sanitizer_${name}.log" 2>&1
  local rc=$?
  echo "== $name ($tool): exit $rc"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|error:" "$out/sanitizer_${name}.log" | tail -6
}
run memcheck_conv_tc memcheck 900 tests/test_gpu_conv_tc.py -k "plain_conv or fused_epilogues or tma_epilogue or wide_input or compensated"
run memcheck_conv memcheck 900 tests/test_gpu_conv.py -k "layer_fp32_direct or networks_vs_golden or head"
run memcheck_metrics memcheck 600 tests/test_gpu_metrics.py -k "not 4k and not full_size"
run memcheck_bicubic memcheck 600 tests/test_gpu_bicubic.py -k "not 4k and not full_size"
run memcheck_tiling memcheck 300 tests/test_gpu_tiling.py
run racecheck_metrics racecheck 600 tests/test_gpu_metrics.py -k "not 4k and not full_size"
run racecheck_bicubic racecheck 600 tests/test_gpu_bicubic.py -k "golden"
run synccheck_metrics synccheck 600 tests/test_gpu_metrics.py -k "not 4k and not full_size"
run synccheck_conv_tc synccheck 900 tests/test_gpu_conv_tc.py -k "plain_conv or wide_input"
