# pipelined mesher (emit_kernel next to mesh_kernel) on the STREAMING pass (every chunk materialised and read)
for cfg in "0 0" "3 1" "3 2" "2 2" "4 1"; do set -- $cfg
  if [ "$1" = "0" ]; then echo "sequential"; timeout 200 python tools/mesh_split.py 2>&1 | python -c "
import json,sys; d=json.load(sys.stdin); k=d['elide0_lean1_summaries0']; print(k, round(k['mesh_kernel_ms']+k['emit_kernel_ms'],3))"
  else echo "pipelined mesh=$1 emit=$2"; BXW_EXP_PIPE_MESH=$1 BXW_EXP_PIPE_EMIT=$2 BXW_PIPE=2 timeout 200 python tools/mesh_split.py 2>&1 | python -c "
import json,sys; d=json.load(sys.stdin); k=d['elide0_lean1_summaries0']; print(k, round(k['mesh_kernel_ms']+k['emit_kernel_ms'],3)); k=d['elide1_lean1_summaries1']; print(k, round(k['mesh_kernel_ms']+k['emit_kernel_ms'],3))"
  fi
done
