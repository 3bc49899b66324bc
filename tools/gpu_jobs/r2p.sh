out=gpurun_out/r2p; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt; nproc >> $out/smi.txt; numactl -H >> $out/smi.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 600 $T bench.py --gpus 8 --steps 20 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err
echo "bench 8 took $SECONDS s" >> $out/bench_8gpu.err; tail -2 $out/bench_8gpu.err
timeout 300 $T tools/link_probe.py --gpus 8 > $out/link_probe_8gpu.json 2> $out/link_probe_8gpu.err
timeout 300 $T tools/bench_config5_bands.py --verify --steps 10 > $out/config5_bands_8gpu.json 2> $out/config5_bands_8gpu.err
cat $out/bench_8gpu.json | head -c 600; echo; cat $out/link_probe_8gpu.json; cat $out/config5_bands_8gpu.json | head -c 800
