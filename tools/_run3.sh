set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/s3c_c2_peq_n8.json 2> gpurun_out/s3c_c2_peq_n8.err
timeout 500 $TR --master-port 29512 bench.py --gpus 8 --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s3c_c4_peq_n8.json 2> gpurun_out/s3c_c4_peq_n8.err
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s3c_c5_gcs_n8.json 2> gpurun_out/s3c_c5_gcs_n8.err
tail -c 600 gpurun_out/s3c_c2_peq_n8.err gpurun_out/s3c_c4_peq_n8.err gpurun_out/s3c_c5_gcs_n8.err
cat gpurun_out/s3c_*.json
