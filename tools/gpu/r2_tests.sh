#!/bin/bash
# GPU test suite (everything marked gpu) + smoke + the default bench line
cd /root/repo
mkdir -p gpurun_out
echo "== pytest -m gpu"
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1 | tail -12
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
