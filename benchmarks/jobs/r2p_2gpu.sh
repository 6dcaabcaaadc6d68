#!/bin/bash
# 2-GPU job: pool threads -> devices from the C++ host, torchrun bench at N = 1 and 2 on the same box
O=gpurun_out/r2p; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
python -m pytest tests/test_integration.py -m gpu -q -k "pool_threads" > $O/pool.log 2>&1
python bench.py --steps 12 --warmup 4 --no-configs --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 12 --warmup 4 > $O/bench_n2.json 2> $O/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 examples/global_lighting.py > $O/lighting.log 2>&1
cat $O/gpus.txt; tail -3 $O/pool.log; python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2p/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("configs",{}).get("batch4",{}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
tail -c 400 $O/bench_ref_n2.json; tail -3 $O/lighting.log
