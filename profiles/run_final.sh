# the command list that produced the final r02 files of this directory (one gpurun call, 1 GPU)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02h_tests.log
python bench.py > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.err; echo bench rc $?
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02h_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02h_ncu_ll.log 2>&1
python profiles/bench_case.py > gpurun_out/r02h_case.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o gpurun_out/prof_r02h python profiles/bench_case.py > gpurun_out/r02h_ncu_full.log 2>&1
tail -3 gpurun_out/r02h_case.log
timeout 100 python profiles/soak_parity.py 60 11 2>&1 | tail -2 | tee gpurun_out/r02h_soak.log
