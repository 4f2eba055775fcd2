#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun; one GPU).  $1 = tag, e.g. r1_final
set -u
tag=${1:-r1}
out=gpurun_out
python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > $out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"extract_csr_kernel|sa_sweep_kernel|index_block_kernel" -s 8 -c 5 \
    -o $out/prof_$tag -f python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > $out/ncu_f_$tag.log 2>&1
python tools/profile_gather.py > $out/gather_plain_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gather_index|index_block_kernel" -s 4 -c 4 \
    -o $out/prof_gather_$tag -f python tools/profile_gather.py > $out/ncu_g_$tag.log 2>&1
tail -3 $out/gather_plain_$tag.log
