set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
nproc >> gpurun_out/r2a/smi.txt; free -g >> gpurun_out/r2a/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench_1024.json 2> gpurun_out/r2a/bench_1024.err
timeout 300 python bench.py --images 128 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2a/bench_128.json 2> gpurun_out/r2a/bench_128.err
timeout 300 python tools/link_probe.py > gpurun_out/r2a/link_probe_1gpu.json 2> gpurun_out/r2a/link_probe.err
tail -5 gpurun_out/r2a/pytest.log
