set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02f_pytest_gpu.log
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r02f_bench_reference_arm.json 2>/dev/null
python bench.py --steps 2 --no-configs --no-e2e --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_steps2.csv python bench.py --steps 2 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:espb_resample_kernel -s 3 -c 1 -f -o gpurun_out/r02f_prof_resample python bench.py --steps 2 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r02f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:espb_transpose_flags -s 3 -c 1 -f -o gpurun_out/r02f_prof_flags python bench.py --steps 2 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r02f_ncu_flags.log 2>&1
python tools/bench_kernels.py > gpurun_out/r02f_hbm_kernels.jsonl 2>/dev/null
python tools/bench_biquad.py >> gpurun_out/r02f_hbm_kernels.jsonl 2>/dev/null
python tools/bench_ni.py > gpurun_out/r02f_ni.jsonl 2>/dev/null
ESPB_NI=0 python tools/bench_ni.py >> gpurun_out/r02f_ni.jsonl 2>/dev/null
python tools/write_probe.py >> gpurun_out/r02f_hbm_kernels.jsonl 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la gpurun_out/r02f_*
