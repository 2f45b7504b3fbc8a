#!/bin/bash
# Round-2 GPU session K (gpurun --gpus 2): the multi-device boundary on two real devices - ecb200_init_multi through the Python
# tests and the plain C client, one context per device in one process, index shards under torchrun (NCCL) with the lincomb
# exchange, bench.py at N = 2 (weak and strong), and ONE call sharded inside the library over 1 / 2 devices
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/s11_box.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 ) | tee $O/s11_pytest_round2.txt
timeout 300 python scripts/two_devices_one_process.py 2>&1 | tail -2 | tee $O/s11_two_devices.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/sharded_2gpu.py 2>&1 | tail -3 | tee $O/s11_sharded_2gpu.txt
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/s11_bench_2gpu_weak.json 2> $O/s11_bench_2gpu_weak.err; echo "weak rc=$?"; cut -c1-300 $O/s11_bench_2gpu_weak.json
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > $O/s11_bench_2gpu_strong.json 2> $O/s11_bench_2gpu_strong.err; echo "strong rc=$?"; cut -c1-300 $O/s11_bench_2gpu_strong.json
timeout 600 python scripts/multi_device_one_call.py k256 22 5 2>&1 | tail -3 | tee $O/s11_one_call_multi_device.txt
