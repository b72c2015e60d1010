#!/bin/bash
# GPU call 26 (last GPU seconds of the round): the default bench line of the final build, and smoke().
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
timeout 75 python bench.py > $O/c26_bench.json 2> $O/c26_bench.err
cut -c1-200 $O/c26_bench.json; tail -1 $O/c26_bench.err
(timeout 20 python -c "import __graft_entry__ as g; g.smoke()") > $O/c26_smoke.log 2>&1
tail -1 $O/c26_smoke.log
