# emit_kernel shapes (experimental builds csrc/libbxw_cuda_exp_v*.so: -DBXW_EMIT_PV / _THREADS / _CTAS)
for v in v1 v2 v3 v4 v5 v6 v7 v8; do
  echo "variant $v"
  BXW_CUDA_LIB=ballxworld_b200/csrc/libbxw_cuda_exp_$v.so python tools/bench_mesh_only.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 mesh', round(d['mesh_kernel_ms'],3), 'emit', round(d['emit_kernel_ms'],3))"
  BXW_CUDA_LIB=ballxworld_b200/csrc/libbxw_cuda_exp_$v.so python tools/mesh_split.py 2>&1 | python -c "import json,sys; d=json.load(sys.stdin); print('terrain', d['elide1_lean1_summaries1'])"
done
