#!/bin/bash
# r2r: full gpu suite, both bench arms, SGM A/Bs
O=gpurun_out/r2r; mkdir -p $O
python benchmarks/sgm_bench.py > $O/sgm_bench.json 2> $O/sgm_bench.err
SMVSB_SGM_NO_F2I=1 python benchmarks/sgm_bench.py > $O/sgm_bench_nof2i.json 2>> $O/sgm_bench.err
timeout 1800 python -m pytest tests -m gpu -q -s --durations=8 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
grep -h "ms_cost_volume" $O/sgm_bench.json $O/sgm_bench_nof2i.json | cut -c1-200; grep -E "passed|failed|rc=" $O/pytest.log | tail -3; grep -h '"job"\|cut_depth_maps' $O/pytest.log | cut -c1-300; tail -1 $O/smoke.log; cat $O/bench.json; cat $O/bench_ref.json | cut -c1-600
