#!/bin/bash
# round 2: parity suite, bench lines of every workload, sanitizer passes on small shapes, ncu launch lists and full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
timeout 200 python tools/timeline.py --head > gpurun_out/timeline.log 2>&1
# ---- compute-sanitizer on small shapes (step / loop / beam / train): memcheck, racecheck, synccheck
SMALL="tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_adam_update_matches_tf_semantics tests/test_gpu_train.py::test_tensor_core_attend_projection_in_training tests/test_gpu_edges.py"
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 8 python -m pytest $SMALL -m gpu -q -x --timeout 400 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -3
done
# ---- ncu: launch list of the decode loop, full captures, training launch list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 320 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-train --pool 2 --profile-run > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:att_wpc -s 30 -c 2 \
    -o gpurun_out/prof_att -f python bench.py --steps 1 --warmup 3 --no-cpu --no-train --pool 1 --profile-run > gpurun_out/ncu_att.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 200 -c 6 \
    -o gpurun_out/prof_lin -f python bench.py --steps 1 --warmup 3 --no-cpu --no-train --pool 1 --profile-run > gpurun_out/ncu_lin.log 2>&1
bash tools/gpu_train_list.sh > /dev/null 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
for f in bench bench_ref bench_cfg3 bench_train1 bench_cfg5; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -n 1 | cut -c1-300; done
tail -n 2 gpurun_out/ncu_list.log gpurun_out/ncu_att.log gpurun_out/ncu_lin.log
