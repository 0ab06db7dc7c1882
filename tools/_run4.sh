set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s3d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3d_tests.log
tail -3 gpurun_out/s3d_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s3d_c2_peq.json 2> gpurun_out/s3d_c2_peq.err
timeout 300 python bench.py --workload c3 --metric jc --no-cpu-baseline > gpurun_out/s3d_c3_jc.json 2> gpurun_out/s3d_c3_jc.err
timeout 300 python bench.py --workload c3 --metric af --no-cpu-baseline > gpurun_out/s3d_c3_af.json 2> gpurun_out/s3d_c3_af.err
timeout 600 python bench.py --workload c4 --genomes 20000 --steps 2 --no-cpu-baseline > gpurun_out/s3d_c4_peq_20k.json 2> gpurun_out/s3d_c4_peq_20k.err
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu-baseline > gpurun_out/s3d_c5_gcs.json 2> gpurun_out/s3d_c5_gcs.err
tail -c 300 gpurun_out/s3d_*.err
for f in gpurun_out/s3d_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("k_blocks_visited_frac"), d["stage_ms_per_step"], d.get("checksum"))
    print(d["e2e"])
except Exception as e: print("ERR", e)
PY
done
