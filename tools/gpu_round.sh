#!/bin/bash
# standard GPU round: parity suite, bench lines for every workload, launch list (ncu), full ncu captures of the
# attention kernel and of the dense kernels, in-kernel timelines.  "lite": tests, bench lines and the two launch lists only;
# "noncu": everything but the profiler passes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
if [ "$1" != "lite" ]; then
timeout 300 python tools/timeline.py --head > gpurun_out/timeline.log 2>&1
timeout 300 python tools/trace.py > gpurun_out/trace.log 2>&1
timeout 300 python tools/trace_loop.py 79 128 96 > gpurun_out/trace_loop.log 2>&1
timeout 300 python tools/att_time.py > gpurun_out/att_time.log 2>&1
fi
if [ "$1" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 320 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 2 --profile-run > gpurun_out/ncu_list.log 2>&1
if [ "$1" != "lite" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_wpc -s 30 -c 2 \
    -o gpurun_out/prof_att -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 --profile-run > gpurun_out/ncu_att.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 200 -c 6 \
    -o gpurun_out/prof_lin -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 --profile-run > gpurun_out/ncu_lin.log 2>&1
fi
bash tools/gpu_train_list.sh > /dev/null 2>&1
fi
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log
for f in bench bench_ref bench_cfg3 bench_train1 bench_cfg5; do echo "== $f"; tail -n 1 gpurun_out/$f.log | cut -c1-400; done
tail -n 3 gpurun_out/ncu_list.log gpurun_out/ncu_att.log gpurun_out/ncu_lin.log
