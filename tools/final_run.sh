#!/bin/bash
# round-end measurement on one B200: full GPU suite, smoke, bench (both arms), ncu launch list, traffic, the full-set capture,
# the next rows, the sanitizer over the new rewrite kernel, the non-ASCII and adversarial sweeps (least important last)
set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.txt 2>&1
timeout 300 python bench.py > gpurun_out/r02_bench_n1_10GiB.json 2> gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_1GiB.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --docs 262144 > gpurun_out/r02_launches_bench.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_1GiB.csv python tools/profile_run.py --docs 262144 --iters 3 > gpurun_out/r02_traffic.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_warp_scan -s 1 -c 1 -o gpurun_out/r02_warp_scan python tools/profile_run.py --docs 65536 > gpurun_out/r02_warp_scan.log 2>&1
timeout 200 python tools/bench_rows.py > gpurun_out/r02_next_rows.json 2> gpurun_out/r02_next_rows.err
timeout 200 python tools/bench_rewrite.py > gpurun_out/r02_rewrite_one_pass.json 2> gpurun_out/r02_rewrite.err
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_format_rows.py -m gpu -x -q -k rewrite > gpurun_out/r02_sanitizer_memcheck_rewrite.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_format_rows.py -m gpu -x -q -k "rewrite and not edges" > gpurun_out/r02_sanitizer_racecheck_rewrite.log 2>&1
timeout 200 python tools/bench_unicode.py > gpurun_out/r02_non_ascii_documents.txt 2>&1
timeout 400 python tools/bench_sweep.py --cell-mib 256 --iters 3 > gpurun_out/r02_sweep_c5_256MiB.json 2> gpurun_out/r02_sweep.err
cat gpurun_out/r02_pytest_gpu.txt gpurun_out/r02_smoke.txt; cut -c1-400 gpurun_out/r02_bench_n1_10GiB.json
