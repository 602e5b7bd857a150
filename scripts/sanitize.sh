#!/bin/bash
# compute-sanitizer over the kernels that depend on zero-on-entry / zero-on-exit global scratch, packed 16-bit
# shared counters, three-stream fork / join and record_stream: the hub-stage tests (folded and exact index
# layouts, segment walker, position windows, heavy-source passes), the cn5 / cn6 aggregation tests (per-link and
# run-grouped kernels) and the graph build / mask tests, at test size.
#   bash scripts/sanitize.sh [outdir]      (run under gpurun; logs -> outdir, default gpurun_out/sanitizer)
out=${1:-gpurun_out/sanitizer}
mkdir -p "$out"
sel='hub_stage_bit_exact or hub_stage_position_windows or hub_stage_heavy or hub_run_segment or run_grouped or cn5_aggregate or cn6_order3 or test_gpu_graph_build or dropadj or out_of_range'
small='not pubmed and not collab_s and not cora-1152-256'
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph_build.py -m gpu -x -q -k "($sel) and ($small)" \
    > "$out/$tool.log" 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' "$out/$tool.log") summaries: $(grep 'ERROR SUMMARY' "$out/$tool.log" | sort | uniq -c | tr '\n' ';')  $(tail -1 "$out/$tool.log")"
done
