K=b200_persistent_speculative_while_while,b200_wide4,b200_sw,b200_wide4_sw
timeout 300 python scripts/kernel_compare.py --kernels $K --batches 8 --repeats 2 --out gpurun_out/r2g_sw_default.json 2>&1 | tail -5
for c in 4 6; do echo "CTAS=$c"; NT_SW_CTAS=$c timeout 300 python scripts/kernel_compare.py --kernels b200_sw,b200_wide4_sw --batches 8 --repeats 2 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'; done
for f in 4 16 24; do echo "FETCH=$f"; NT_SW_FETCH=$f timeout 300 python scripts/kernel_compare.py --kernels b200_sw,b200_wide4_sw --batches 8 --repeats 2 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'; done
echo "SMEM=4"; NT_SW_SMEM=4 timeout 300 python scripts/kernel_compare.py --kernels b200_sw,b200_wide4_sw --batches 8 --repeats 2 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'
echo "NODE_EXIT=0"; NT_SW_NODE_EXIT=0 timeout 300 python scripts/kernel_compare.py --kernels b200_sw,b200_wide4_sw --batches 8 --repeats 2 2>&1 | grep -o '^b200[a-z_0-9]* {"primary": [0-9.]*, "AO": [0-9.]*, "diffuse": [0-9.]*'
