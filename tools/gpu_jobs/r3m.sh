out=gpurun_out/r3m; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err
timeout 300 python bench.py --no-cpu --no-e2e > $out/bench_1gpu_same_box.json 2> $out/bench_1gpu.err
cat $out/bench_8gpu.json | head -c 300; echo
