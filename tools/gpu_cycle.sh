#!/bin/bash
# One GPU round trip (run under gpurun): parity tests, smoke, the bench line, and the ncu evidence for
# profiles/ (launch list + --set full of one timed step).  Usage: bash tools/gpu_cycle.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; tail -3 $O/pytest_$TAG.log
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
python bench.py > $O/bench_$TAG.log 2> $O/bench_$TAG.err; tail -c 5000 $O/bench_$TAG.log; tail -3 $O/bench_$TAG.err
python tools/time_stages.py > $O/stages_$TAG.log 2>&1; cat $O/stages_$TAG.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu1_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:logmel_kernel|sepconv" -s 48 -c 4 -f \
    -o $O/prof_$TAG python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu2_$TAG.log 2>&1
echo "ncu rc=$?"
