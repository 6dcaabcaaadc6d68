#!/bin/bash
# r2ae: 2-GPU job with the kernels of the end of round 2: torchrun bench at N = 1 and 2 on the same box, as the driver launches it
O=gpurun_out/r2ae; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
python bench.py --no-configs --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err
python -m pytest tests/test_integration.py -m gpu -q -k "pool_threads" > $O/pool.log 2>&1
cat $O/gpus.txt; tail -2 $O/pool.log; python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2ae/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["two_host_threads_per_gpu"]["value"], d.get("configs",{}).get("batch4",{}).get("value"), d.get("configs",{}).get("shading",{}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
tail -3 $O/bench_n2.err
