TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29531 tools/pcie_probe_ranks.py > gpurun_out/pcie_ranks_n8.json 2> gpurun_out/pcie_ranks_n8.err; echo "probe rc=$?"; cat gpurun_out/pcie_ranks_n8.json
$TR --master-port 29532 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench numa rc=$?"
EXO_BENCH_NUMA=0 $TR --master-port 29533 bench.py --gpus 8 > gpurun_out/bench_n8_nonuma.json 2> gpurun_out/bench_n8_nonuma.err; echo "bench no-numa rc=$?"
python - <<'PY'
import json
for f in ("bench_n8", "bench_n8_nonuma"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms %.4f frac %.3f e2e %.4g numa %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("numa_binding")))
        for k, v in (d.get("multi_gpu") or {}).items():
            print("   ", k, "ms %.4f 1gpu %.4f eff %.3f coll_us %.1f" % (v["ms_per_step"], v["ms_per_step_1gpu"], v["strong_scaling_efficiency"], v["collective_us"]))
    except Exception as ex:
        print(f, "failed", ex)
PY
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head
