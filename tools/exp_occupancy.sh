# mesh_kernel CTAs per SM (BXW_EXP_MESH_CTAS) on the bench terrain. (The emit_kernel half of this sweep, quoted in
# profiles/r02_summary.md as "emit v4", was run with an env knob that the final one-CTA-per-SM emit_kernel no longer has; its
# shape is now chosen at compile time, see tools/exp_emit_variants.sh.)
for m in 4 3 2; do echo "mesh=$m"; BXW_EXP_MESH_CTAS=$m timeout 100 python tools/mesh_split.py 2>&1 | tr -d "\n " | cut -c 1-95; echo; done
