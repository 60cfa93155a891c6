set -x
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.err
timeout 300 python tools/profile_step.py > gpurun_out/plain_final4.log 2>&1; tail -1 gpurun_out/plain_final4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:opus:: -s 224 -c 1200 --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py > gpurun_out/ncu_l4.log 2>&1; tail -1 gpurun_out/ncu_l4.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_bf16 -s 135 -c 4 -f -o gpurun_out/prof_prefill_gemm_final2 python tools/profile_step.py > gpurun_out/ncu_f4.log 2>&1; tail -1 gpurun_out/ncu_f4.log
ncu -i gpurun_out/prof_prefill_gemm_final2.ncu-rep --page raw --csv > gpurun_out/prof_prefill_gemm_final2.csv 2>/dev/null; wc -c gpurun_out/prof_prefill_gemm_final2.csv
