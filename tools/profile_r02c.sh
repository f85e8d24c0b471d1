# end of round 2: the update kernel with the cost epilogue and the retuned cost kernels (tools/bench_next.py launches them on cfg3)
mkdir -p gpurun_out/prof3
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_tree_update_cost --launch-skip 3 --launch-count 1 -f -o /tmp/r02_uc python tools/bench_next.py --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_uc.ncu-rep --tag update_cost
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_cost_block --launch-skip 3 --launch-count 1 -f -o /tmp/r02_cb python tools/bench_next.py --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_cb.ncu-rep --tag cost_block
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_cost_gradient --launch-skip 3 --launch-count 1 -f -o /tmp/r02_cg python tools/bench_next.py --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_cg.ncu-rep --tag cost_gradient
cp profiles/r02_update_cost_* profiles/r02_cost_block_* profiles/r02_cost_gradient_* gpurun_out/prof3/
ls -la gpurun_out/prof3
