#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/gpu_configure_time.py 2>&1 | tee gpurun_out/r02k_configure_time.log
timeout 600 python scripts/gpu_configure_time.py 2>&1 | tee -a gpurun_out/r02k_configure_time.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -q -m gpu -k "vector_error or emitter_sampling_only or reference_source_goldens" 2>&1 | tail -8
