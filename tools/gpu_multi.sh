#!/bin/bash
# multi-GPU round: bash tools/gpu_multi.sh N tag [pytest]
N=$1; tag=$2; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > $out/topo_$tag.txt 2>&1
if [ "$3" = "pytest" ]; then
  ( time python -m pytest tests -x -q -m gpu ) > $out/pytest_$tag.log 2>&1; tail -4 $out/pytest_$tag.log
fi
$TR tools/pcie_probe.py > $out/pcie_multi_$tag.json 2> $out/pcie_multi_$tag.err; cut -c1-400 $out/pcie_multi_$tag.json
$TR bench.py --gpus $N --steps 200 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err; tail -3 $out/bench_$tag.err
python - <<PY
import json
d=json.load(open("$out/bench_$tag.json"))
for k in ("value","scaling","ms_per_step","weak","sustained","e2e","e2e_ceiling","e2e_reference_api","checks","issue_roofline"):
    print(k, json.dumps(d.get(k))[:500])
PY
