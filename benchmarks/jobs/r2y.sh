#!/bin/bash
# r2y: full gpu suite (CPU sides of the full-size tests computed on the box), both bench arms with all configs
O=gpurun_out/r2y; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
tail -14 $O/pytest.log; cat $O/bench_ref.json | cut -c1-600; cat $O/bench.json | cut -c1-1500
