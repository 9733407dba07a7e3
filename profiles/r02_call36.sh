#!/bin/bash
# Round 2, GPU call 36 (1 GPU): harness tests after counting the launches of the gated explicit-flag pair.
python -m pytest tests/test_harness.py tests/test_bench_contract.py -q -m gpu --maxfail=5 2>&1 | tail -3
