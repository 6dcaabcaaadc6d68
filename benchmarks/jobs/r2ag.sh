#!/bin/bash
# r2ag: what the driver runs at round end -- the gpu suite, smoke(), both bench arms
O=gpurun_out/r2ag; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x --durations=6 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
tail -12 $O/pytest.log; tail -2 $O/smoke.log; cut -c1-300 $O/bench_ref.json; cut -c1-900 $O/bench.json
