# emit_kernel shapes: threads per CTA (one CTA per SM) and vertices per pass are compile-time constants. Build variants with
#   cd ballxworld_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static \
#     --fmad=false -DBXW_EMIT_THREADS=512 -DBXW_EMIT_PV=64 -shared -o libbxw_cuda_exp_v1.so *.cu tables.cpp
# (v1, v2, ...) and run this script on the GPU: hashed shapes (configs[1]) and the bench terrain per variant. The sweep that chose
# 384 threads x 160 vertices is tabulated in profiles/r02_summary.md (during the sweep CTAs per SM and the number of staging
# buffers were macros as well).
for lib in ballxworld_b200/csrc/libbxw_cuda_exp_v*.so; do
  echo "variant $lib"
  BXW_CUDA_LIB=$lib python tools/bench_mesh_only.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 mesh', round(d['mesh_kernel_ms'],3), 'emit', round(d['emit_kernel_ms'],3))"
  BXW_CUDA_LIB=$lib python tools/mesh_split.py 2>&1 | python -c "import json,sys; d=json.load(sys.stdin); print('terrain', d['elide1_lean1_summaries1'])"
done
