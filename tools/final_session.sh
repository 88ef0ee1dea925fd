O=gpurun_out; T=r02v
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.txt 2>&1
python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.txt 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err ) 2> $O/${T}_bench.time
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err ) 2> $O/${T}_bench_reference.time
export KJB_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python tools/profile_frames.py --scene atrium --ircache --rtr --taa --spatial 2 --frames 8 > $O/${T}_launches.log 2>&1
tail -2 $O/${T}_smoke.txt; tail -3 $O/${T}_gpu_tests.txt; cat $O/${T}_bench.time $O/${T}_bench_reference.time; python -c "
import json;d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e']['ms_per_step'],d['fast_math'].get('ms_per_step'));[print('  ',e['config']['workload'],round(e['ms_per_step'],3),round(e['e2e']['ms_per_step'],3)) for e in d['configs']]"
