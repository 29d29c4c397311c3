#!/bin/bash
# usage: tools/launches.sh <tag> [bench args...]   -> gpurun_out/launches_<tag>.csv  (run under gpurun)
tag=$1; shift
timeout -s KILL 90 python bench.py "$@" > gpurun_out/plain_$tag.log 2>&1 && \
timeout -s KILL 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches_$tag.csv python bench.py "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -c 200 gpurun_out/ncu_$tag.log
