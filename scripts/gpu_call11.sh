#!/bin/bash
# GPU call 11: the whole GPU suite on the final build, smoke(), one default bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
mkdir -p $O
(time timeout 1800 python -m pytest tests -m "gpu and not slow" -x -q) > $O/c11_pytest.log 2>&1
echo "pytest rc=$?" >> $O/c11_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > $O/c11_smoke.log 2>&1
echo "smoke rc=$?" >> $O/c11_smoke.log
timeout 600 python bench.py --steps 8 --warmup 5 --no-cpu-baseline > $O/c11_bench.json 2> $O/c11_bench.err
tail -4 $O/c11_pytest.log; tail -2 $O/c11_smoke.log; cut -c1-200 $O/c11_bench.json
