# dev tool (run under gpurun, one GPU): the ncu captures of round 2 that are not the trace kernel's
set -x
mkdir -p gpurun_out
# 1. --set full over the non-trace kernel families (second pass of scripts/profile_misc.py)
REGEX='regex:raygen|count_hits|ray_key|ray_aabb|ray_reorder|radix|scan_chained|layout|basic|collapse|pack_kernel|sah|wide4_convert|hlbvh_top|cluster|morton|topology|finalize|emit|tri_box|word'
timeout 900 ncu --set full --clock-control none --import-source on -k "$REGEX" -o gpurun_out/r2k_misc_full -f python scripts/profile_misc.py > gpurun_out/r2k_misc_full.log 2>&1
ncu -i gpurun_out/r2k_misc_full.ncu-rep --page raw --csv > gpurun_out/r2k_misc_full_raw.csv 2>/dev/null
python scripts/ncu_summarise.py gpurun_out/r2k_misc_full_raw.csv gpurun_out/r2k_misc_full_summary.md > /dev/null
ls -la gpurun_out/r2k_misc_full*
rm -f gpurun_out/r2k_misc_full.ncu-rep            # gpurun brings back at most 64 MiB; the raw CSV export holds every metric of the capture
gzip -f gpurun_out/r2k_misc_full_raw.csv
# 2. what binds the trace kernels once the BVH no longer fits L2: 10.5 M-triangle scene (839 MB BVH)
timeout 900 python scripts/ncu_binding.py --kernel b200_auto --scene sanmiguel --hlbvh-bits 4 --out gpurun_out/r2k_binding_b200_auto_sanmiguel.json > gpurun_out/r2k_binding_sanmiguel.log 2>&1
tail -c 400 gpurun_out/r2k_binding_sanmiguel.log
# 3. launch list of the default bench command (kernel share of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_bench_launches.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/r2k_bench_profile.log 2>&1
tail -3 gpurun_out/r2k_bench_profile.log
