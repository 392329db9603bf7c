set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2_gpu_tests.log 2>&1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_ref.err
python bench_icp.py > gpurun_out/r2_bench_icp.json 2>/dev/null
python bench_frontier.py > gpurun_out/r2_bench_frontier.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_tiled_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-merge > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_home_raycast|k_home_count|k_home_scatter|k_tile_plan|k_home_resolve" -s 20 -c 5 -f -o gpurun_out/r2_tiled python bench.py --steps 2 --warmup 3 --no-cpu --no-merge > gpurun_out/ncu_t.log 2>&1
STEPS=4 ncu --set full --clock-control none --import-source on -k regex:"k_home_raycast" -s 8 -c 1 -f -o gpurun_out/r2_fused python tools/profile_band_step.py > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_icp_loop" -c 1 -f -o gpurun_out/r2_icp python bench_icp.py > gpurun_out/ncu_i.log 2>&1
tail -3 gpurun_out/r2_gpu_tests.log
