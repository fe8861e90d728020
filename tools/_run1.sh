export C4_FZ_TIMEOUT_S=120
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_selfplay.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --no-cpu --no-extras > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python tools/show_bench.py gpurun_out/bench_default.json
