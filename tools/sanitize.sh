#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests: the hand-rolled mbarrier / bulk-copy rings
# (covariance, source-model, fused, activation, ISS kernels), the cluster / DSMEM NMF kernel, CUDA-graph replay with cached
# pointer signatures and the multi-stream pipelined job.  Logs go to gpurun_out/ (copied to profiles/ when clean).
#   tools/sanitize.sh [seconds per tool]
cd "$(dirname "$0")/.."
LIMIT=${1:-420}
OUT=gpurun_out
mkdir -p $OUT
SEL_BROAD='gauss_ilrma_golden or auxiva_golden or tilrma_golden or fastmnmf_golden or nmf_golden or weighted_covariance or ip_update or consistent'
SEL_RING='ilrma_ip_power_d2 or ilrma_iss_power_d2 or auxiva_laplace_ip or fused_source_model or cluster_kernel or graph_replay_equals_eager_loop or pipelined_waveform'
run_as() {  # tool, log name, selection, files...
  tool=$1; name=$2; sel=$3; shift 3
  log=$OUT/sanitizer_${name}.log
  echo "== compute-sanitizer --tool $tool : pytest -k \"$sel\" $*" > $log
  timeout $LIMIT compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest "$@" -x -q -m gpu -k "$sel" -p no:cacheprovider >> $log 2>&1
  echo "== exit code $?  (124 = stopped at the ${LIMIT}s limit, 99 = sanitizer errors)" >> $log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" $log | tail -5
}
run() { t=$1; shift; run_as $t $t "$@"; }
run memcheck "$SEL_BROAD" tests/test_gpu_parity.py tests/test_gpu_mnmf.py tests/test_gpu_nmf.py
run_as memcheck memcheck_ring "$SEL_RING" tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_nmf.py tests/test_gpu_batch.py tests/test_gpu_stft.py
run racecheck "$SEL_RING" tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_nmf.py tests/test_gpu_batch.py
run synccheck "$SEL_RING" tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_nmf.py tests/test_gpu_batch.py
