#!/bin/bash
# round-end verification (+ an optional sweep of the leg's split-K heuristics)
mkdir -p gpurun_out
export OVN_DEBUG_SYNC=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_all.log
unset OVN_DEBUG_SYNC
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
: > gpurun_out/sweep.log
for cfg in "3 1" "2 1" "1 1" "3 2" "1 2"; do
  set -- $cfg
  TAG="min_slabs=$1 fill=$2" OVN_LEG_MIN_SLABS=$1 OVN_LEG_FILL=$2 timeout 200 python tools/time_heads.py >> gpurun_out/sweep.log 2>&1
done
echo done
