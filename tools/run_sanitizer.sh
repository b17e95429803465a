#!/bin/bash
# compute-sanitizer over a few tiny GPU tests (one attention, one GEMM, one state-machine case).
#   tools/run_sanitizer.sh <tool: memcheck|racecheck|synccheck|initcheck> <log> [pytest node ids...]
# Round-1 lesson: the first `import torch` on a fresh box takes ~1 min, longer than the sanitizer's default launch
# time-out ("No attachable process found") -> page the image in first and raise --launch-timeout.
set -u
TOOL=${1:-memcheck}; LOG=${2:-gpurun_out/sanitizer_$TOOL.log}; shift 2 || true
if [ $# -eq 0 ]; then
  set -- "tests/test_gpu_attention.py::test_attention_steady_shapes_vs_oracle[64-15-5-15-2-2-5-2]" \
         "tests/test_gpu_attention.py::test_attention_prefill_causal_vs_oracle[300-2-2-3-2]" \
         "tests/test_gpu_state_machine.py::test_device_state_machine_matches_reference_trace[tiny_bf16_w5n3g3]" \
         "tests/test_gpu_gemm.py::test_gemm_matches_fp32_reference[120-1000-1024-64-2]" \
         "tests/test_gpu_layer_ops.py" \
         "tests/test_gpu_sampling_device.py::test_eos_stops_sampling_and_window_is_filtered" \
         "tests/test_gpu_attention.py::test_attention_lp_shapes_vs_oracle[2]"
fi
python -c "import torch; torch.zeros(1).cuda()" >/dev/null 2>&1
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
timeout ${SAN_TIMEOUT:-900} "$SAN" --tool "$TOOL" --target-processes all --launch-timeout 600 \
    --error-exitcode 97 --print-limit 40 ${SAN_EXTRA:-} \
    python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > "$LOG" 2>&1
rc=$?
echo "sanitizer $TOOL rc=$rc" >> "$LOG"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" "$LOG" | tail -5
exit 0
