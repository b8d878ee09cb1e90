#!/bin/bash
# A/B of prebuilt library variants (variants/*.so) on the benchmark's own workloads: usage  tools/ab_bench.sh <workload> <variant>...
export PYTHONPATH=$PWD
w=$1; shift
for v in "$@"; do
  echo "== $v ($w)"
  ZC_B200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --workload $w --no-extra --no-cpu-baseline --selfplay-games 0 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench value %.4g e2e %.4g kernel_ms %.3f'%(d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms']))"
done
