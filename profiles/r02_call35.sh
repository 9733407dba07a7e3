#!/bin/bash
# Round 2, GPU call 35 (1 GPU): segmented-scan parity after the dispatch clean-up (one explicit-flag instantiation for every type).
python -m pytest tests/test_parity_gpu.py tests/test_harness.py -q -m gpu -k "segscan or flags or config3 or scan" --maxfail=5 2>&1 | tail -2
