# ncu --set full of the chain kernel variants on cfg2a (profiles launches 4-5 of each run); text summaries only come back
NCU="ncu --set full --clock-control none --import-source on --kernel-name regex:exo_chain --launch-skip 4 --launch-count 1"
run() { # name, env..., kernel
  name=$1; kern=$2; shift 2
  env "$@" $NCU -f -o /tmp/$name python tools/bench_configs.py --only ${CFG:-cfg2a} --kernel $kern --reps 5 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/ncu_$name.txt 2>&1
  python tools/ncu_source.py /tmp/$name.ncu-rep --top 40 > gpurun_out/sass_$name.txt 2>&1
}
run stg 2 A=1
run tma64_t4 7 EXO_CHAIN_TMA_TILES=4
run tma64_t1 7 EXO_CHAIN_TMA_TILES=1
run tma128_t4 8 EXO_CHAIN_TMA_TILES=4
ls -la gpurun_out/
