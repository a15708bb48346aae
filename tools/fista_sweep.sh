for v in 0 1 0 1; do
  echo -n "l2_prefetch=$v: "
  DECOMP_GEMM_L2_PREFETCH=$v python bench.py --workload fista --no-cpu --steps 100 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
done
