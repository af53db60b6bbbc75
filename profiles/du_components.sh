# diagnostic: k_bconv_du with its loads / MMAs / stores switched off one at a time (results meaningless, timing only)
for d in 0 1 2 4 3 7; do
  echo -n "ADN_DU_DBG=$d  "; ADN_DU_DBG=$d python bench.py --steps 20 --warmup 3 --no-cpu --no-graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v['ms_per_step']*1000,1) for k,v in d['kernels'].items() if k in ('k_bconv_du',)})"
done
