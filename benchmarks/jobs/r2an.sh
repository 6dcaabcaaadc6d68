#!/bin/bash
# r2an: the whole gpu suite and smoke() on the final commit of round 2
O=gpurun_out/r2an; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
tail -4 $O/pytest.log; tail -1 $O/smoke.log
