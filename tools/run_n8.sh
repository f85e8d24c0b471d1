TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29541 tools/comm_check.py 2> gpurun_out/comm_check8.err | tee gpurun_out/comm_check_peer_n8.json; echo "peer rc=${PIPESTATUS[0]}"
timeout 300 $TR --master-port 29532 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
for k, v in (d.get("multi_gpu") or {}).items():
    print("   ", k, "ms %.4f nocoll %.4f 1gpu %.4f eff %.3f coll_us %.1f" % (v["ms_per_step"], v["ms_without_collective"], v["ms_per_step_1gpu"], v["strong_scaling_efficiency"], v["collective_us"]), v.get("collective_path"), {a: b for a, b in v.items() if "err" in a})
PY
