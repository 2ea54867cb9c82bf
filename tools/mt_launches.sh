#!/bin/bash
# ncu launch list of one MuseTalk B=8 step (graph replay) -> gpurun_out/<tag>_launches_musetalk.csv ; run on the GPU box
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1500 --csv --log-file gpurun_out/${tag}_launches_musetalk.csv \
    python bench_musetalk.py --steps 1 --warmup 1 > /dev/null 2>&1
tail -1 gpurun_out/${tag}_launches_musetalk.csv | cut -c1-120
