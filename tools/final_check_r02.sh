mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_gpu_zz_bus_paired.py -m gpu -q -p no:cacheprovider > gpurun_out/r02h_pytest_zz.log 2>&1; echo "zz rc=$?" >> gpurun_out/r02h_pytest_zz.log )
tail -5 gpurun_out/r02h_pytest_zz.log
( timeout 360 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo "bench rc=$?" >> gpurun_out/bench_r02_final_n1.err )
tail -c 600 gpurun_out/bench_r02_final_n1.json; tail -3 gpurun_out/bench_r02_final_n1.err
( timeout 240 python -m pytest tests/test_gpu_zscale.py -m gpu -q -p no:cacheprovider > gpurun_out/r02h_pytest_zscale.log 2>&1; echo "zscale rc=$?" >> gpurun_out/r02h_pytest_zscale.log )
tail -4 gpurun_out/r02h_pytest_zscale.log
( export KB_BENCH_NO_RANDBENCH=1 KB_BENCH_NO_CLI=1; timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r02_final.log 2>&1; echo "ncu rc=$?" )
