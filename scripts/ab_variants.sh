#!/bin/bash
# A/B on the GPU box: bench stage times for the in-tree library and every variants/libcm_*.so (see build_variant.sh)
for lib in "" variants/libcm_*.so; do
  [ -n "$lib" ] && export CM_LIB_PATH=$PWD/$lib || unset CM_LIB_PATH
  python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batch_of_frames or full_size" 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-latency --sustained-seconds 0 ${AB_BENCH_ARGS} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('${lib:-in-tree}', round(d['value']), d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"
done
