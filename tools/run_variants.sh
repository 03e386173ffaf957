#!/bin/bash
# on the GPU box: time each built variant (see tools/variants.sh); usage: tools/run_variants.sh name...
cd "$(dirname "$0")/.."
for v in "$@"; do
  cp proxtv_b200/variants/lib_$v.so proxtv_b200/libproxtv_b200.so
  echo "=== $v"
  python tools/kbench.py 4096 4096 f64 auto 2>&1 | head -2
  python tools/drprof.py 4096 chunked 2>&1 | tail -2
  python tools/kbench.py 2048 8192 f32 auto 2>&1 | head -2
done
