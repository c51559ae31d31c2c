# dev tool: tile size / direction grid of the coherent ray-generation order (nt_raygen_set_order(1)) on the bench frame, kernel b200_auto
run() { echo "== $*"; env "$@" python scripts/kernel_compare.py --kernels b200_auto --batches 24 --repeats 3 --raygen-order ${ORDER:-1} 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'; }
ORDER=0 run NT_X=0
for r in 1 2 4; do for c in 4 8 16; do run NT_RAYGEN_TILE_R=$r NT_RAYGEN_CELLS=$c; done; done
