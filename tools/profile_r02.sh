# Runs on the GPU box: ncu captures of the headline kernel (bench.py's own loop, caches NOT flushed between replays so that the
# DRAM bytes are those of back-to-back launches) and of the generated tree kernel (cfg4); text summaries go to gpurun_out/.
set -x
mkdir -p gpurun_out/prof
# 1. launch list of the default bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/prof/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/prof/ncu_launches.log 2>&1
# 2. headline kernel, full set, back-to-back launches of the timed loop
ncu --set full --clock-control none --cache-control none --import-source on --kernel-name regex:exo_chain_tma --launch-skip 8 --launch-count 2 -f -o /tmp/r02_chain \
    python bench.py --steps 20 --warmup 3 > gpurun_out/prof/ncu_chain.log 2>&1
python tools/make_profiles.py --round r02 --launches gpurun_out/prof/launches.csv --rep /tmp/r02_chain.ncu-rep --tag chain --traffic-json
# 3. generated tree kernel (cfg4)
ncu --set full --clock-control none --cache-control none --import-source on --kernel-name regex:exo_tree --launch-skip 4 --launch-count 2 -f -o /tmp/r02_tree \
    python tools/bench_configs.py --only cfg4 --reps 5 > gpurun_out/prof/ncu_tree.log 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_tree.ncu-rep --tag tree
# 4. the other chain configurations and the sphere kernels with the default cache control (one launch each)
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_ --launch-skip 3 --launch-count 1 -f -o /tmp/r02_cfg3 \
    python tools/bench_configs.py --only cfg3 --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_cfg3.ncu-rep --tag cfg3
cp profiles/r02_* profiles/traffic.json gpurun_out/prof/
ls -la gpurun_out/prof
