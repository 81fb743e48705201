# emit_kernel CTAs per SM (BXW_EXP_EMIT_CTAS): hashed shapes (configs[1]) and the bench terrain
for n in 1 2 3; do
  echo "emit CTAs/SM $n"
  BXW_EXP_EMIT_CTAS=$n python tools/bench_mesh_only.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 mesh', round(d['mesh_kernel_ms'],3), 'emit', round(d['emit_kernel_ms'],3))"
  BXW_EXP_EMIT_CTAS=$n python tools/mesh_split.py 2>&1 | python -c "import json,sys; d=json.load(sys.stdin); print('terrain', d['elide1_lean1_summaries1'], 'streaming', d['elide0_lean1_summaries0'])"
done
