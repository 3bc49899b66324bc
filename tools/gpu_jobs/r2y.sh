out=gpurun_out/r2y; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q --timeout=150 -k "float64_stage or either_chunk_size" > $out/pytest_new.log 2>&1; echo "pytest rc=$?" >> $out/pytest_new.log; tail -3 $out/pytest_new.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err
echo "bench 8 took $SECONDS s" >> $out/bench_8gpu.err; tail -2 $out/bench_8gpu.err
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T4 bench.py --gpus 4 --steps 20 --warmup 3 > $out/bench_4gpu.json 2> $out/bench_4gpu.err
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $T2 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_2gpu.json 2> $out/bench_2gpu.err
timeout 300 $T tools/bench_config5_bands.py --verify --steps 10 > $out/config5_bands_8gpu.json 2> $out/config5_bands_8gpu.err
cat $out/bench_8gpu.json | head -c 400; echo
