set -x
timeout 900 python -m pytest tests/test_gpu_held_forward.py -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_gpu_deferred.py tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_lazy.json 2> gpurun_out/bench_lazy.err; tail -c 600 gpurun_out/bench_lazy.err; cat gpurun_out/bench_lazy.json | cut -c1-300
VKHEL_LAZY_FORWARD=0 python bench.py --steps 20 --warmup 3 2>/dev/null | cut -c1-200
