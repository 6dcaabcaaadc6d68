#!/bin/bash
# r2l: final PCG kernel timing, full gpu suite, both bench arms
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
python benchmarks/cg_probe.py new > $O/cg_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
tail -3 $O/cg_probe.log; tail -5 $O/pytest.log; cat $O/bench.json; cat $O/bench_ref.json
