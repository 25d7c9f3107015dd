#!/bin/bash
# tools/ab.sh name1 name2 ...: bench the variant libraries under variants/ (run under gpurun)
for v in "$@"; do
  ARPEGGIO_CUDA_LIB=$PWD/variants/lib_$v.so python bench.py --steps 200 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$v', 'ms/step %.4f' % d['ms_per_step'], 'grid %.4f search %.4f classify %.4f' % (r['grid_build_ms'], r['search_ms'], r['classify_ms']), 'e2e ms %.3f' % d['e2e']['ms_per_step'])"
done
