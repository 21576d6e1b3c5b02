#!/bin/bash
# tools/profile_round.sh <tag> -- the ncu evidence of a round, run under gpurun on ONE B200 (recipe:
# /opt/skills/guides/B200_PROFILING.md).  Writes gpurun_out/<tag>_*: the launch list of a short bench.py run and one
# --set full capture of each kernel family.  Every profiled command is first run plain (ncu only after exit 0).
set -u
tag=${1:-r02}
out=gpurun_out
B="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --queue-blocks 0 --program-blocks 0"
$B > $out/${tag}_plain_bench.json 2> $out/${tag}_plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cn_block_kernel|index_map_kernel|cn_bytes_kernel|inflate_tiles_kernel|cn_deflate_fused_kernel|deflate_tiles_kernel|ship_strip_kernel" -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_bench.log 2>&1
python tools/kbench.py --planes 9 --rows-per-cta 0 --once > $out/${tag}_plain_k9.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cn_block_kernel -c 1 -o $out/${tag}_cn_block_9 python tools/kbench.py --planes 9 --rows-per-cta 0 --once > $out/${tag}_ncu_k9.log 2>&1
python tools/kbench.py --planes 1 --rows-per-cta 0 --once > $out/${tag}_plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cn_block_kernel -c 1 -o $out/${tag}_cn_block_1 python tools/kbench.py --planes 1 --rows-per-cta 0 --once > $out/${tag}_ncu_k1.log 2>&1
python tools/one_block.py --reps 1 > $out/${tag}_plain_block.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:inflate_tiles_kernel -c 1 -o $out/${tag}_inflate python tools/one_block.py --reps 1 > $out/${tag}_ncu_inflate.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_deflate_fused_kernel -s 4 -c 1 -o $out/${tag}_fused python tools/one_block.py --reps 1 > $out/${tag}_ncu_fused.log 2>&1
for f in $out/${tag}_ncu_*.log; do tail -n 2 "$f"; done
