set -x
timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 330 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 230 -c 6 -o gpurun_out/r1c_gemm_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
