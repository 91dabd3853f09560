mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config train --steps 5 --warmup 2 > gpurun_out/r2h_train_2gpu.json 2> gpurun_out/r2h_train_2gpu.err; tail -c 900 gpurun_out/r2h_train_2gpu.json; tail -3 gpurun_out/r2h_train_2gpu.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1200 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err; head -c 400 gpurun_out/r2h_bench_2gpu.json
