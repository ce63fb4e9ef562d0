set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest8.log 2>&1; tail -3 gpurun_out/r2v_pytest8.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-drop-in > gpurun_out/r2v_n1.json 2> gpurun_out/r2v_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_n$N.json 2> gpurun_out/r2v_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --batch-per-gpu 64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2v_n8_b64.json 2> gpurun_out/r2v_n8_b64.err
python bench.py --gpus 1 --batch-per-gpu 128 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-drop-in > gpurun_out/r2v_n1_b128.json 2> gpurun_out/r2v_n1_b128.err
python - <<'PY'
import json
for n in ("n1","n2","n4","n8","n8_b64","n1_b128"):
    try:
        d=json.load(open(f"gpurun_out/r2v_{n}.json")); print(n, round(d["value"],1), d["n_gpus"], round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["config"].get("batch_per_gpu"))
    except Exception as e: print(n, "ERR", e)
PY
