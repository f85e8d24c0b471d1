# Multi-GPU checks on one box (run under `gpurun --gpus N`): NGPU=2|4|8 bash tools/run_multi_gpu.sh
#   1. tools/comm_check.py: collectives against seeds through the peer windows, then through NCCL (EXO_COMM_PEER=0)
#   2. tools/pcie_probe_ranks.py: what the node's host complex moves when every rank copies one e2e step's bytes at once
#   3. bench.py --gpus N: the contract line with its multi_gpu section
# Every command runs under its own timeout: a wedged collective must not hold the box.
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 120 $TR --master-port 29541 tools/comm_check.py 2> gpurun_out/comm_check.err | tee gpurun_out/comm_check_peer_n$N.json; echo "peer rc=${PIPESTATUS[0]}"
EXO_COMM_PEER=0 timeout 120 $TR --master-port 29542 tools/comm_check.py 2>> gpurun_out/comm_check.err | tee gpurun_out/comm_check_nccl_n$N.json; echo "nccl rc=${PIPESTATUS[0]}"
timeout 120 $TR --master-port 29543 tools/pcie_probe_ranks.py 2> gpurun_out/pcie_ranks.err | tee gpurun_out/pcie_ranks_n$N.json; echo "probe rc=${PIPESTATUS[0]}"
timeout 300 $TR --master-port 29544 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("value %.4g ms %.4f frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
for k, v in (d.get("multi_gpu") or {}).items():
    print("   ", k, "ms %.4f nocoll %.4f 1gpu %.4f eff %.3f coll_us %.1f" % (v["ms_per_step"], v["ms_without_collective"], v["ms_per_step_1gpu"], v["strong_scaling_efficiency"], v["collective_us"]), v.get("collective_path"))
PY
