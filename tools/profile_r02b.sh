mkdir -p gpurun_out/prof2
python tools/cfg3_step_probe.py 512 | tee gpurun_out/prof2/cfg3_step_512.txt
python tools/cfg3_step_probe.py 4096 | tee -a gpurun_out/prof2/cfg3_step_512.txt
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 12 --launch-count 8 --csv python tools/cfg3_step_probe.py 512 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>10 and r[0]!='ID': print('  ncu:', r[4][:70], r[8], r[-1], 'ns')
" | tee -a gpurun_out/prof2/cfg3_step_512.txt
# generated fused cost-gradient kernel of the whole-body problem and the park kernel: full sets
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_tree_costgrad --launch-skip 3 --launch-count 1 -f -o /tmp/r02_tcg python tools/bench_next.py --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_tcg.ncu-rep --tag tree_costgrad
ncu --set full --clock-control none --import-source on --kernel-name regex:"exo_tree_park|exo_spheres" --launch-skip 4 --launch-count 2 -f -o /tmp/r02_sph python tools/bench_configs.py --only cfg5 --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_sph.ncu-rep --tag spheres
ncu --set full --clock-control none --import-source on --kernel-name regex:exo_chain_costgrad --launch-skip 3 --launch-count 1 -f -o /tmp/r02_ccg python tools/bench_next.py --reps 3 > /dev/null 2>&1
python tools/make_profiles.py --round r02 --rep /tmp/r02_ccg.ncu-rep --tag costgrad
cp profiles/r02_tree_costgrad_* profiles/r02_spheres_* profiles/r02_costgrad_* gpurun_out/prof2/
ls gpurun_out/prof2
