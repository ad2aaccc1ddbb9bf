#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck and synccheck over the hand-written kernels at the
# parity tests' small shapes -- both tcgen05 conv kernels (implicit GEMM incl. the split-precision and paired-tile paths,
# halo-resident conv incl. folded / composite / few-CTA ring-wrap cases) and the small kernels.  Run on the GPU box:
#     bash tools/sanitize.sh            # logs under gpurun_out/sanitizer_*.log (copy the summaries to profiles/)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='few_ctas or split_precision or layer0_split or epilogue_noise or prelu_residual or fc_as_conv or finalize_and_affine or torgb or noise_project or match_vote or (conv_matches_torch and (2-64-64-16-1-3 or 2-64-64-112-2-3 or 2-64-128-56-2-1)) or (upconv_phases and 64-32-16) or (halo_conv_matches_torch) or (halo_upconv_matches_torch) or (composite_matches_torch) or (blur_act_stats and (2-32-64 or 1-512-8))'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_${tool}.log \
      python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$SEL" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "exit $?"
  tail -3 gpurun_out/sanitizer_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_${tool}.log | tail -5
done
