set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s3j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3j_tests.log
tail -3 gpurun_out/s3j_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s3j_c2_peq.json 2> gpurun_out/s3j_c2_peq.err
PC_HOST_THREADS=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/s3j_c2_peq_t1.json 2> gpurun_out/s3j_c2_peq_t1.err
timeout 600 python bench.py --workload c4 --genomes 12000 --steps 2 --no-cpu-baseline > gpurun_out/s3j_c4_peq_12k.json 2> gpurun_out/s3j_c4_peq_12k.err
PC_HOST_THREADS=1 timeout 600 python bench.py --workload c4 --genomes 12000 --steps 2 --no-cpu-baseline > gpurun_out/s3j_c4_peq_12k_t1.json 2> gpurun_out/s3j_c4_peq_12k_t1.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s3j_smoke.log 2>&1
tail -2 gpurun_out/s3j_smoke.log
tail -c 300 gpurun_out/s3j_*.err
for f in gpurun_out/s3j_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d.get("checksum"))
    print(d["e2e"]["ms_per_step"], d["e2e"]["stage_ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
