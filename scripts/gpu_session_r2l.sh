#!/bin/bash
# Round-2 GPU session L (gpurun --gpus 8): STRONG scaling of the BASELINE configs named "on 8xB200" - one 2^22-row verify batch
# (secp256k1, P-256) and one 2^20-row P-384 / SM2 P*k batch cut into 8 index shards, one process per GPU under torchrun - and ONE
# ecb200_ecdsa_verify call sharded inside the library over 1 / 2 / 4 / 8 devices of one process
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/s12_box.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > $O/s12_bench_8gpu_strong_k256.json 2> $O/s12_k256.err; echo "k256 rc=$?"; cut -c1-300 $O/s12_bench_8gpu_strong_k256.json
timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --curve p256 > $O/s12_bench_8gpu_strong_p256.json 2> $O/s12_p256.err; echo "p256 rc=$?"; cut -c1-300 $O/s12_bench_8gpu_strong_p256.json
timeout 400 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --op mul_var --curve p384 > $O/s12_bench_8gpu_strong_p384_mul.json 2> $O/s12_p384.err; echo "p384 rc=$?"; cut -c1-300 $O/s12_bench_8gpu_strong_p384_mul.json
timeout 400 $TR --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --distinct-keys > $O/s12_bench_8gpu_strong_k256_distinct.json 2> $O/s12_k256d.err; echo "k256 distinct rc=$?"; cut -c1-300 $O/s12_bench_8gpu_strong_k256_distinct.json
timeout 400 python scripts/multi_device_one_call.py k256 22 5 2>&1 | tail -5 | tee $O/s12_one_call_multi_device.txt
