set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 2 --workload c4 --genomes 4000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s3f_c4_n2.json 2> gpurun_out/s3f_c4_n2.err
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --workload c5 --genomes 20000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/s3f_c5_n2.json 2> gpurun_out/s3f_c5_n2.err
tail -c 1500 gpurun_out/s3f_c4_n2.err gpurun_out/s3f_c5_n2.err
cat gpurun_out/s3f_c4_n2.json gpurun_out/s3f_c5_n2.json
