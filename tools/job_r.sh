#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py tests/test_abi_and_host.py -m gpu -x -q -k "flux or abi or symbol or golden" 2>&1 | tail -6
