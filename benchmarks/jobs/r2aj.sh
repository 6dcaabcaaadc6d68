#!/bin/bash
# r2aj: the bench line of the final library (SGM with the 16-byte-load sum / WTA kernel)
O=gpurun_out/r2aj; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2aj/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_construct"]["frac"], d["roofline_construct"]["pipe_frac"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step", v.get("ms_per_run")), v.get("ms_cost_paths_wta"), v.get("roofline",{}).get("frac"))
PY
tail -3 $O/bench.err
