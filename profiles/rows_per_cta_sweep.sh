for r in 14 15 16 17 19 20; do
  echo "rows_per_cta=$r"; ADN_ROWS_PER_CTA=$r python bench.py --steps 30 --warmup 3 --no-cpu --no-graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', round(d['ms_per_step'],4), {k: round(v['ms_per_step']*1000,1) for k,v in d['kernels'].items() if k in ('k_fconv','k_bconv_du')})"
done
