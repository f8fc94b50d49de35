#!/bin/bash
# CPU-side check of the whole tree (no GPU needed): compile the CUDA library for sm_100a, the oracle, oracle/_ref (when the reference tree is
# present), the Rcpp adapters of src/ against oracle/shim and the vendor comparator; then the `not gpu` test suite.  On a B200: add `-m gpu`.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as e; e.build(); print('build ok')"
python -m pytest tests/ -x -q -m "not gpu"
if python -c "import torch, sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
    python -m pytest tests/ -x -q -m gpu
    python -c "import __graft_entry__ as e; e.smoke()"
fi
