TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1"
timeout 90 $TR --master-port 29541 tools/comm_check.py 2> gpurun_out/comm_check.err | tee gpurun_out/comm_check_peer.json; echo "peer rc=${PIPESTATUS[0]}"
EXO_COMM_PEER=0 timeout 90 $TR --master-port 29542 tools/comm_check.py 2>> gpurun_out/comm_check.err | tee gpurun_out/comm_check_nccl.json; echo "nccl rc=${PIPESTATUS[0]}"
tail -5 gpurun_out/comm_check.err
timeout 240 $TR --master-port 29543 bench.py --gpus ${NGPU:-2} > gpurun_out/bench_nX.json 2> gpurun_out/bench_nX.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_nX.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
for k, v in (d.get("multi_gpu") or {}).items():
    print("   ", k, "ms %.4f 1gpu %.4f eff %.3f coll_us %.1f" % (v["ms_per_step"], v["ms_per_step_1gpu"], v["strong_scaling_efficiency"], v["collective_us"]), {a: b for a, b in v.items() if "err" in a})
PY
