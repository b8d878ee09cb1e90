#!/bin/bash
# A/B timing of prebuilt library variants (variants/*.so): kernel time and instruction-cache hit rates of the chess fused search
export PYTHONPATH=$PWD
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__inst_executed.sum,gpu__time_duration.sum
for v in "$@"; do
  echo "== $v"
  ZC_B200_LIB=$PWD/variants/$v.so timeout 300 python tools/quick_bench_chess.py 16384 800 2>&1 | grep "sims/s" | tail -2
  ZC_B200_LIB=$PWD/variants/$v.so timeout 300 ncu --metrics $M --clock-control none -k regex:k_search_fused -c 1 -s 2 --csv python tools/quick_bench_chess.py 16384 800 2>/dev/null | grep -E '^"0"' | awk -F'","' '{print "   ", $13, $15}'
done
