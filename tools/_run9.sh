set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s3i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3i_tests.log
tail -3 gpurun_out/s3i_tests.log
timeout 300 python bench.py --workload c3 --metric jc --no-cpu-baseline > gpurun_out/s3i_c3_jc.json 2> gpurun_out/s3i_c3_jc.err
timeout 300 python bench.py --workload c3 --metric pocp --no-cpu-baseline > gpurun_out/s3i_c3_pocp.json 2> gpurun_out/s3i_c3_pocp.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s3i_c2_peq.json 2> gpurun_out/s3i_c2_peq.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:i8_contract -s 3 -c 1 -o gpurun_out/s3i_prof_i8 python bench.py --workload c3 --metric jc --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s3i_ncu_i8.log 2>&1
tail -c 300 gpurun_out/s3i_*.err
for f in gpurun_out/s3i_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d.get("checksum"))
    print(d["e2e"]["ms_per_step"], d["e2e"]["stage_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
