for m in 4 3 2; do for e in 4 3 2; do echo "mesh=$m emit=$e"; BXW_EXP_MESH_CTAS=$m BXW_EXP_EMIT_CTAS=$e timeout 100 python tools/mesh_split.py 2>&1 | tr -d "\n " | cut -c 1-95; echo; done; done
for e in 4 3 2; do echo "cfg2 emit=$e"; BXW_EXP_EMIT_CTAS=$e timeout 100 python tools/bench_mesh_only.py | cut -c 150-300; done
