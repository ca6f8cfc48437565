#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "atm_ or hbv_single or gr4j_tree" 2>&1 | tail -6
