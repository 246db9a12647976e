#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "kat or blob_kzg_proof_vectors or batch_vectors or synthetic_batch_64 or canonicity or helper" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -6 gpurun_out/pytest_quick.log
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1
tail -5 gpurun_out/probe.log | head -3
