mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_default.json
CMD="python bench.py --steps 2 --warmup 3 --framesets 128 --no-cpu-baseline --no-e2e --no-workloads --no-shim"
$CMD > gpurun_out/ncu_plain.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fuse_stream -s 4 -c 1 -o gpurun_out/r02_stream_c $CMD > gpurun_out/ncu_f.log 2>&1; echo "ncu rc=$?"
