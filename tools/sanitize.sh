#!/bin/bash
# compute-sanitizer memcheck over one small case of every kernel family (run on a GPU box; output under gpurun_out/).
set -u
mkdir -p gpurun_out
run() { echo "=== $*" >> gpurun_out/sanitizer.log; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 "$@" >> gpurun_out/sanitizer.log 2>&1; echo "exit code $?" >> gpurun_out/sanitizer.log; }
run python -m pytest tests/test_gpu_dense_block.py -q -m gpu -x -k "test_dense16_kernel_both_sides_vs_fp64 and (333-257 or 640-500 or 129-33)"
run python -m pytest tests/test_gpu_dense_block.py -q -m gpu -x -k "test_spmm_dense_block_matches_plain_csr_and_fp64 and (333-257 or 640-500)"
run python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "(grad_w and (777 or 3001 or 2000-16-804)) or single_product"
run python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_spmm_all_outputs or test_block_aggregate_fwd_bwd_vs_oracle"
run python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "test_blocked_split_with_fused_column_sums and (5-16 or 1234-132 or 300-36)"
run python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "test_linear_tc_forward_backward and dense16 and not 20000 and not 5003"
run python -m pytest tests/test_gpu_peer.py -q -m gpu -x -k "test_local_ranks and (1 or 2)"
grep -E "^===|exit code|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer.log
