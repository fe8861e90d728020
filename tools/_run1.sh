export C4_FZ_TIMEOUT_S=120
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_default.json
