timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r1b.err
python tools/show_bench.py gpurun_out/bench_r1b.json 2>&1 | head -40
