# compute-sanitizer over the hot path at small shapes (hand-rolled mbarrier / TMEM / TMA code: SURVEY 5):
#   gpurun --timeout 1500 -- 'bash tools/sanitize_round.sh 2>&1 | tail -30'
# memcheck: out-of-bounds / misaligned global, shared and local accesses; racecheck: shared-memory hazards between threads
# of a CTA; synccheck: invalid barrier use.  Logs -> gpurun_out/sanitize_*.log (copied to profiles/ when clean).
set -x
mkdir -p gpurun_out
SEL='smoke or (golden and not fullres) or tail_golden or bins_head or mix_weight or (ws_kernels and (cfg0 or cfg1 or cfg2 or cfg3)) or (test_summary_tc and (cfg0 or cfg1 or cfg2)) or median or rotation'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" \
      -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -4
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitize_smoke_memcheck.log 2>&1; echo "smoke memcheck rc=$?"
