python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generated_tree or valkyrie or cfg4 or humanoid or layouts" 2>&1 | tail -3
run() { echo "$* : "; env "$@" python tools/bench_configs.py --only ${CFGS:-cfg4} --kernel 9 --reps 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('   ',d['config'],d.get('kernel'),'%.1f us (best %.1f) frac %.3f'%(d['ms_median']*1e3,d['ms_best']*1e3,d['frac_of_measured_hbm']) if 'ms_median' in d else d)
    except Exception: print(l.rstrip()[:300])
"; }
CFGS=cfg2a,cfg2b,cfg3,cfg4
run EXO_TREE_TILE=32
run EXO_TREE_TILE=64
CFGS=cfg4
run EXO_TREE_TILE=64 EXO_TREE_WPC=2
run EXO_TREE_TILE=64 EXO_TREE_WPC=8
run EXO_TREE_TILE=64 EXO_TREE_CAP=12
run EXO_TREE_TILE=64 EXO_TREE_ZC=2
run EXO_TREE_TILE=64 EXO_TREE_ZC=6
run EXO_TREE_TILE=64 EXO_TREE_AHEAD=12
