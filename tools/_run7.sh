set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s3g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3g_tests.log
tail -3 gpurun_out/s3g_tests.log
timeout 300 python bench.py --workload c3 --metric jc --no-cpu-baseline > gpurun_out/s3g_c3_jc.json 2> gpurun_out/s3g_c3_jc.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s3g_c2_peq.json 2> gpurun_out/s3g_c2_peq.err
tail -c 300 gpurun_out/s3g_*.err
nproc
for f in gpurun_out/s3g_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d.get("checksum"))
    print(d["e2e"])
except Exception as e: print("ERR", e)
PY
done
