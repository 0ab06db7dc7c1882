set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s3b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3b_tests.log
tail -5 gpurun_out/s3b_tests.log
timeout 600 python bench.py > gpurun_out/s3b_c2_peq.json 2> gpurun_out/s3b_c2_peq.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s3b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s3b_launches_c2_peq.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s3b_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nw_ -s 60 -c 12 -o gpurun_out/s3b_prof_dp python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s3b_ncu_dp.log 2>&1
tail -3 gpurun_out/s3b_c2_peq.json
