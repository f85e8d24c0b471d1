python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generated_tree or valkyrie or cfg4 or humanoid" 2>&1 | tail -4
run() { echo -n "$* : "; env "$@" python tools/bench_configs.py --only cfg4 --reps 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['config'],d.get('kernel'),'%.1f us (best %.1f) frac %.3f'%(d['ms_median']*1e3,d['ms_best']*1e3,d['frac_of_measured_hbm']) if 'ms_median' in d else d)
    except Exception: print(l.rstrip()[:300])
"; }
run EXO_TREE_AHEAD=0
run EXO_TREE_AHEAD=6
run EXO_TREE_AHEAD=12
run EXO_TREE_CAP=6
run EXO_TREE_CAP=12
run EXO_TREE_CAP=18
run EXO_TREE_CAP=12 EXO_TREE_WPC=2
run EXO_TREE_CAP=12 EXO_TREE_ZC=2
run EXO_TREE_CAP=12 EXO_TREE_AHEAD=12
run EXO_TREE_CAP=6 EXO_TREE_ZC=2 EXO_TREE_WPC=2
