#!/bin/bash
# development helper: one compressed-in / compressed-out block per library build under variants/
for f in variants/libgcn10cuda_*.so; do
    n=$(basename $f .so); n=${n#libgcn10cuda_}
    GCN10_CUDA_LIB=$PWD/$f python tools/one_block.py --reps ${REPS:-6} --planes ${PLANES:-9} 2>&1 | tail -1 | python -c "
import sys, json, statistics as st
d = json.loads(sys.stdin.read())[1:]
print('$n', 'block_ms', round(st.median(x['block_ms'] for x in d), 3), 'inflate', round(st.median(x['inflate_ms'] for x in d), 3), 'fused_sum', round(st.median(x['fused_ms_sum_over_strips'] for x in d), 2), 'bytes', d[0]['out_bytes'])
"
done
