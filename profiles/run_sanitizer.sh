#!/bin/bash
# compute-sanitizer passes over the hand-rolled synchronisation (tcgen05 + mbarrier pipelines, PDL overlap, the
# device-wide spin barrier of the persistent step) on small shapes; run on the GPU box:
#   gpurun -- 'bash profiles/run_sanitizer.sh'
# Every pass writes gpurun_out/sanitizer_<tool>_<case>.log; the ERROR SUMMARY lines are collected into
# gpurun_out/sanitizer_summary.txt (copied to profiles/r02_sanitizer_summary.txt).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SUM=gpurun_out/sanitizer_summary.txt
: > $SUM
run() {   # tool case args...
  local tool=$1 name=$2; shift 2
  local log=gpurun_out/sanitizer_${tool}_${name}.log
  local t0=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --launch-timeout 120 "$@" > $log 2>&1
  local rc=$?
  echo "[$tool $name] rc=$rc $(( $(date +%s) - t0 )) s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $SUM
}
export USE_GRAPH=0       # eager launches: the sanitizer attributes every access to its own kernel launch
# batch 12: tcgen05 GEMM chain (tc_small_gemm, PDL), fused decode attention, sampler, prefill GEMMs + flash attention,
# then the vocoder (tc_conv_gemm / tc_halo_conv / tc_halo_bulk / tc_pair_conv, conv_post)
for tool in memcheck racecheck synccheck; do
  VITS_TOKENS=6 run $tool b12_vits python tests/gpu_prof_small.py 12 3 vits
done
# batch 2: persistent decode step (cooperative launch, device-wide barrier, cross-barrier prefetch)
for tool in memcheck racecheck synccheck; do
  run $tool b2_persistent python tests/gpu_prof_small.py 2 3
done
cat $SUM
