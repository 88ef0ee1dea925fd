#!/bin/bash
# One GPU session (run under gpurun): tests, the driver's two bench commands, a launch list and an ncu capture summarised ON the box
# (the .ncu-rep of a multi-kernel `--set full` capture exceeds what gpurun copies back; the csv pages do not).
#   tools/gpu_session.sh <tag> [ncu kernel regex] [skip] [count]
TAG=${1:-r02x}; RE=${2:-"k_"}; SKIP=${3:-250}; CNT=${4:-42}   # default: every kernel of one steady-state frame
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.txt 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ) 2> $O/${TAG}_bench.time
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err ) 2> $O/${TAG}_bench_reference.time
export KJB_NO_GRAPH=1   # profile directly launched kernels
PF="python tools/profile_frames.py --scene atrium --ircache --rtr --taa --spatial 2 --frames 8"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv $PF > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -o /tmp/${TAG}_full $PF > $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_source.csv 2>/dev/null
head -c 20000 /tmp/${TAG}_source.csv > $O/${TAG}_source_head.csv
python tools/ncu_hot_lines.py /tmp/${TAG}_source.csv 60 > $O/${TAG}_hot_lines.txt 2>&1
ls -la $O /tmp/${TAG}_full.ncu-rep
